"""The wrappers the reference's entry points put around the model, on the real kernels:
  * `DistributedDataParallel(model, device_ids=[LOCAL_RANK])` (scripts/train/train_model_builder.py:75-78): the train-mode
    forward is one autograd node over all parameters, so DDP's gradient hooks fire and `.grad` is populated;
  * `nn.SyncBatchNorm.convert_sync_batchnorm(model)` (:86-91): SyncBatchNorm modules are executed by the same fused BN
    kernels (cross-rank sums when the world is larger than one; here world = 1 must equal plain BatchNorm bit for bit);
  * `model.half()` + `imgs.half()` (val.py:331-334 / train_utils.py:436-444): fp16 parameters and inputs are accepted."""
import os

import pytest
import torch
import torch.nn as nn

from _parity import errs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nccl_world_of_one():
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield dist
    if dist.is_initialized():
        dist.destroy_process_group()


def _loss(outs):
    return sum((o.float() ** 2).mean() for o in outs)


def test_ddp_and_syncbn_wrappers(nccl_world_of_one):
    from copy import deepcopy

    from ayolov2_b200 import synth

    torch.cuda.set_device(0)
    base = synth.build_model("yolov5n", seed=0).cuda().train()
    x = torch.rand(4, 3, 128, 128, device="cuda")
    # plain
    plain = deepcopy(base)
    _loss(plain(x)).backward()
    g_plain = [p.grad.clone() for p in plain.parameters()]
    # SyncBatchNorm-converted (train_model_builder.py:86-91), wrapped in DDP (:75-78)
    sync = nn.SyncBatchNorm.convert_sync_batchnorm(deepcopy(base))
    assert any(isinstance(m, nn.SyncBatchNorm) for m in sync.modules())
    ddp = nn.parallel.DistributedDataParallel(sync, device_ids=[0])
    _loss(ddp(x)).backward()
    g_ddp = [p.grad for p in ddp.module.parameters()]
    assert all(g is not None for g in g_ddp)
    num = sum(float((a - b).norm() ** 2) for a, b in zip(g_ddp, g_plain)) ** 0.5
    den = sum(float(b.norm() ** 2) for b in g_plain) ** 0.5
    assert num / den < 1e-3, num / den  # same kernels, same statistics: only the atomics' summation order differs
    # the BatchNorm running statistics advanced identically
    for (n1, b1), (n2, b2) in zip(plain.named_buffers(), ddp.module.named_buffers()):
        if b1.dtype.is_floating_point:
            assert torch.allclose(b1, b2, rtol=1e-5, atol=1e-6), n1
    # eval mode through the SyncBatchNorm modules (running statistics folded like BatchNorm2d)
    ev_plain, _ = plain.eval()(x)
    ev_sync, _ = ddp.module.eval()(x)
    assert torch.allclose(ev_plain, ev_sync, rtol=1e-3, atol=1e-3)


def test_half_model_and_inputs():
    from ayolov2_b200 import synth
    from oracle import yolo_oracle

    model = synth.build_model("yolov5n", seed=1)
    x = torch.rand(2, 3, 160, 128, generator=torch.Generator().manual_seed(3))
    want, _ = yolo_oracle.forward(model, x)
    m16 = model.cuda().half().eval()          # val.py:331-334
    got, raw = m16(x.cuda().half())           # train_utils.py:440-444: `imgs.half() if self.half else imgs`
    assert got.shape == want.shape and len(raw) == 3
    e = errs(got[..., :4], want[..., :4])
    assert e["max_norm"] < 1e-2 and e["rel_l2"] < 1e-2, e
    assert float((got[..., 4:].float().cpu() - want[..., 4:]).abs().max()) < 1e-2
