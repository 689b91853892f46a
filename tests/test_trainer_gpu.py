"""ayolov2_b200.trainer.TrainStep (the drop-in of YoloTrainer.training_step, scripts/train/yolo_trainer.py:289-358) against a
shadow built from the reference's own recipe in plain PyTorch: `torch.optim.SGD(nesterov)` with the three parameter groups of
`_init_optimizer` (:140-168), the warm-up ramps of `warmup` (:194-221), gradient accumulation (:331-338) and the ModelEMA
update (torch_utils.py:405-416), all fed with the SAME gradients (the engine's flat gradient buffer). What is compared is
therefore the schedule + the fused optimizer/EMA kernel, parameter by parameter."""
from copy import deepcopy

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

HYP = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, lrf=0.1, momentum=0.937,
           weight_decay=5e-4, warmup_epochs=3.0, warmup_momentum=0.8, warmup_bias_lr=0.1,
           optimizer_params=dict(lr=0.01, momentum=0.937, nesterov=True))


def _targets(bs, seed):
    g = torch.Generator().manual_seed(seed)
    n = 3 * bs
    t = torch.zeros(n, 6)
    t[:, 0] = torch.randint(0, bs, (n,), generator=g).float()
    t[:, 1] = torch.randint(0, 80, (n,), generator=g).float()
    t[:, 2:4] = 0.1 + 0.8 * torch.rand(n, 2, generator=g)
    t[:, 4:6] = 0.05 + 0.3 * torch.rand(n, 2, generator=g)
    return t


def _reference_optimizer(model, hyp, batch_size, accumulate):
    """yolo_trainer.py:140-168 in plain torch."""
    pg0, pg1, pg2 = [], [], []
    for _, v in model.named_modules():
        if hasattr(v, "bias") and isinstance(v.bias, torch.Tensor):
            pg2.append(v.bias)
        if isinstance(v, nn.BatchNorm2d):
            pg0.append(v.weight)
        elif hasattr(v, "weight") and isinstance(v.weight, torch.Tensor):
            pg1.append(v.weight)
    opt = torch.optim.SGD(pg0, **hyp["optimizer_params"])
    opt.add_param_group({"params": pg1, "weight_decay": hyp["weight_decay"] * batch_size * accumulate / 64})
    opt.add_param_group({"params": pg2})
    for g in opt.param_groups:
        g["initial_lr"] = g["lr"]
    return opt


@pytest.mark.parametrize("batch_size,first_batch", [(16, 0), (16, 600)])
def test_training_step_schedule_and_fused_optimizer(batch_size, first_batch):
    from ayolov2_b200 import synth
    from ayolov2_b200.trainer import TrainStep, lr_function

    model = synth.build_model("yolov5n", seed=0)
    shadow = deepcopy(model).cuda().train()
    ema = deepcopy(shadow).eval()
    nb, epochs = 50, 30
    ts = TrainStep(model, HYP, batch_size=batch_size, batches_per_epoch=nb, epochs=epochs, img_size=128)
    accumulate0 = max(round(64 / batch_size), 1)
    opt = _reference_optimizer(shadow, HYP, batch_size, accumulate0)
    num_warmups = max(round(HYP["warmup_epochs"] * nb), 1e3)
    names = [n for n, _ in model.named_parameters()]
    sp = dict(shadow.named_parameters())
    updates, steps_done, accumulate = 0, 0, accumulate0
    for i in range(7):
        bi = first_batch + i
        g = torch.Generator().manual_seed(100 + i)
        imgs = torch.randint(0, 256, (batch_size, 3, 128, 128), generator=g, dtype=torch.uint8)
        loss = ts.training_step((imgs, _targets(batch_size, i), None, None), bi, 0)
        assert torch.isfinite(loss).all()
        eng = ts._engine()
        # ---- shadow: the reference recipe on the same gradients
        ni = bi
        if ni <= num_warmups:  # warmup(), yolo_trainer.py:194-221
            xs = [0, num_warmups]
            accumulate = max(1, np.interp(ni, xs, [1, 64 / batch_size]).round())
            for j, x in enumerate(opt.param_groups):
                x["lr"] = np.interp(ni, xs, [HYP["warmup_bias_lr"] if j == 2 else 0.0, x["initial_lr"] * lr_function(0, epochs, HYP["lrf"])])
                x["momentum"] = np.interp(ni, xs, [HYP["warmup_momentum"], HYP["momentum"]])
        assert ts.accumulate == accumulate
        flat = eng.last_grad_flat
        for name, p, o in zip(names, model.parameters(), eng.pg_offsets):
            gslice = flat[o:o + p.numel()].view(p.shape).clone()
            sp[name].grad = gslice if sp[name].grad is None else sp[name].grad + gslice  # backward() accumulates into .grad
        if ni % accumulate == 0:
            opt.step()
            opt.zero_grad()
            updates += 1
            d = 0.9999 * (1 - np.exp(-updates / 2000))
            msd = shadow.state_dict()
            for k, v in ema.state_dict().items():
                if v.dtype.is_floating_point:
                    v *= d
                    v += (1.0 - d) * msd[k].detach()
            steps_done += 1
        # ---- compare every parameter and the EMA parameters
        for name, p in model.named_parameters():
            assert torch.allclose(p.data, sp[name].data, rtol=1e-5, atol=1e-7), (i, name)
        for name, v in ts.ema_state_dict().items():
            if name in sp:
                assert torch.allclose(v, dict(ema.named_parameters())[name].data, rtol=1e-5, atol=1e-7), (i, name)
        # BN running statistics differ between model (updated by its forward) and shadow (never run): sync for the EMA check
        for (kb, vb), (_, sb) in zip(model.named_buffers(), shadow.named_buffers()):
            sb.copy_(vb)
    assert steps_done >= 2 and (first_batch == 0 or steps_done < 7)  # the accumulation case skips optimizer steps


def test_backward_twice_without_forward_is_refused():
    """the engine keeps one set of activations; a stale backward must raise, not return wrong gradients."""
    from ayolov2_b200 import synth

    model = synth.build_model("yolov5n", seed=0).cuda().train()
    x = torch.rand(2, 3, 64, 64, device="cuda")
    out1 = model(x)
    out2 = model(x)
    with pytest.raises(RuntimeError, match="ONE set of activations"):
        sum(o.sum() for o in out1).backward()
    sum(o.sum() for o in out2).backward()
    assert all(p.grad is not None for p in model.parameters() if p.requires_grad)


def test_multi_scale_and_bucket_plan():
    from ayolov2_b200 import synth
    from ayolov2_b200.trainer import TrainStep

    model = synth.build_model("yolov5n", seed=0)
    ts = TrainStep(model, HYP, batch_size=64, batches_per_epoch=10, epochs=3, img_size=128, multi_scale=True)
    shapes = set()
    for i in range(4):
        imgs = torch.randint(0, 256, (4, 3, 128, 128), dtype=torch.uint8)
        loss = ts.training_step((imgs, _targets(4, i), None, None), i, 0)
        assert torch.isfinite(loss).all()
        shapes.add((ts._engine().H, ts._engine().W))
    assert all(h % 32 == 0 and 64 <= h <= 224 for h, _ in shapes)
    # bucket plan: contiguous, covering, in execution order (last layers first)
    eng = ts._engine()
    chunks = eng.plan_grad_buckets(4)
    assert 2 <= len(chunks) <= 4
    assert chunks[0][3] == eng.pg_flat.numel() and chunks[-1][2] == 0 and chunks[-1][0] == 0 and chunks[0][1] == len(eng.bwd)
    for a, b in zip(chunks[:-1], chunks[1:]):
        assert a[2] == b[3] and a[0] == b[1]
    # the bucketed backward computes the same gradient as the single-graph backward
    x = torch.rand(4, 3, eng.H, eng.W, device="cuda")
    model.train()
    outs = model(x)
    sum((o * o).sum() for o in outs).backward()
    g_bucketed = eng.last_grad_flat.clone()
    del eng.bwd_chunks
    eng._gstate = {k: v for k, v in eng._gstate.items() if not k.startswith("bwd")}
    outs = model(x)
    sum((o * o).sum() for o in outs).backward()
    # (split-K weight gradients accumulate with fp32 atomics: run-to-run order differs, so compare in norm)
    rel = float((eng.last_grad_flat - g_bucketed).norm() / g_bucketed.norm())
    assert rel < 1e-3, rel


def test_gradient_buffers_carry_nothing_between_steps():
    """The activation-gradient buffers are static and never zero-filled: every slice is overwritten by its first writer of
    the step (train_engine._GradSite). A step's gradient must therefore not depend on what earlier steps left behind --
    compare the 4th backward of one engine (eager, captured, replayed before it, on a 10x larger input) with the first
    backward of a fresh engine on the same batch."""
    from ayolov2_b200 import synth

    def flat_grad(model, xs):
        model.train()
        for x in xs:
            model.zero_grad(set_to_none=True)
            outs = model(x)
            sum((o * o).sum() for o in outs).backward()
        return torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).clone()

    g = torch.Generator(device="cuda").manual_seed(3)
    x_big = 10.0 * torch.rand(2, 3, 96, 128, device="cuda", generator=g)
    x = torch.rand(2, 3, 96, 128, device="cuda", generator=g)
    used = flat_grad(synth.build_model("yolov5s", seed=0).cuda(), [x_big, x_big, x_big, x])
    fresh = flat_grad(synth.build_model("yolov5s", seed=0).cuda(), [x])
    rel = float((used - fresh).norm() / fresh.norm())
    assert rel < 2e-3, rel  # split-K weight gradients accumulate with fp32 atomics: run-to-run order differs
