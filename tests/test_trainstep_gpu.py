"""End-to-end parity of one training step (BASELINE.json configs[2] at a small size): train-mode forward with batch
statistics -> ComputeLoss -> backward through every layer, CUDA path vs the CPU fp32 oracle (restated operators +
loss oracle + autograd). bf16 activations/gradients: tolerances are relative L2 per tensor and a global cosine."""
from copy import deepcopy

import pytest
import torch

pytestmark = pytest.mark.gpu

HYP = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, label_smoothing=0.0)


def _targets(bs, n, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(n, 6)
    t[:, 0] = torch.randint(0, bs, (n,), generator=g).float()
    t[:, 1] = torch.randint(0, 80, (n,), generator=g).float()
    t[:, 2:4] = 0.1 + 0.8 * torch.rand(n, 2, generator=g)
    t[:, 4:6] = 0.05 + 0.4 * torch.rand(n, 2, generator=g)
    return t


def _grad_stats(m, ref):
    dots = n1s = n2s = 0.0
    worst = (0.0, "")
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        g1, g2 = p.grad.detach().cpu().double(), q.grad.detach().cpu().double()
        dots += float((g1 * g2).sum())
        n1s += float((g1 * g1).sum())
        n2s += float((g2 * g2).sum())
        rel = float((g1 - g2).norm() / (g2.norm() + 1e-12))
        if g2.norm() > 1e-6 and rel > worst[0]:
            worst = (rel, n)
    return dots / (n1s ** 0.5 * n2s ** 0.5), (n1s / n2s) ** 0.5, worst


ACT = {"activation": "SiLU"}
ANCH = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]


def mini_cfg(v5: bool):
    """A 16-layer YOLOv5-shaped graph with every operator of the full model (6x6 stem or Focus, stride-2 convs, C3 with
    and without shortcut, SPPF or SPP, UpSample, Concat with fan-out, 3-level YOLOHead). A random-init batch-normalised
    network is chaotic in depth (gradient noise grows exponentially with the number of BN layers); at this depth the
    bf16 rounding alone leaves a gradient cosine of ~0.975 against exact fp32, so an exact backward is measurable."""
    stem = [-1, 1, "Focus", [64, 3], ACT] if v5 else [-1, 1, "Conv", [64, 6, 2, 2], ACT]
    pool = [-1, 1, "SPP", [512, [5, 9, 13]], ACT] if v5 else [-1, 1, "SPPF", [512, 5], ACT]
    return {"input_size": [256, 256], "input_channel": 3, "depth_multiple": 0.33, "width_multiple": 0.5, "anchors": ANCH,
            "n_classes": 80, "activation": "SiLU",
            "backbone": [stem, [-1, 1, "Conv", [128, 3, 2], ACT], [-1, 3, "C3", [128], ACT], [-1, 1, "Conv", [256, 3, 2], ACT],
                         [-1, 3, "C3", [256], ACT], [-1, 1, "Conv", [512, 3, 2], ACT], pool, [-1, 1, "Conv", [256, 1, 1], ACT],
                         [-1, 1, "UpSample", [None, 2]], [[-1, 4], 1, "Concat", [1]], [-1, 3, "C3", [256, False], ACT],
                         [-1, 1, "Conv", [256, 3, 2], ACT], [[-1, 7], 1, "Concat", [1]], [-1, 3, "C3", [512, False], ACT],
                         [-1, 1, "Conv", [512, 3, 2], ACT], [-1, 3, "C3", [512, False], ACT]],
            "head": [[[10, 13, 15], 1, "YOLOHead", [80, ANCH]]]}


def _build(name):
    from ayolov2_b200 import synth

    if name.startswith("mini"):
        import kindle

        torch.manual_seed(0)
        return kindle.YOLOModel(mini_cfg(name == "mini_v5"), init_bias=True)
    return synth.build_model(name, seed=0)


@pytest.mark.parametrize("name,hw,bs,min_cos", [("mini_v6", (256, 256), 4, 0.975), ("mini_v5", (256, 256), 4, 0.975),
                                                ("yolov5s", (160, 192), 2, 0.80)])
def test_backward_matches_oracle_for_fixed_upstream_gradient(name, hw, bs, min_cos):
    """Backward through every layer (BN batch statistics, SiLU, conv dgrad/wgrad, shortcut, concat, upsample, SPP(F),
    head) for a FIXED gradient on the three head outputs: sum_i <pred_i, G_i>. This isolates the model backward from
    the loss, whose objectness/class gradients sigma(x) - t are exponentially sensitive to the (bf16-noisy) logits."""
    from oracle import yolo_oracle

    base = _build(name)
    x = torch.rand((bs, 3, *hw), generator=torch.Generator().manual_seed(3))
    # exact fp32 oracle (only to report how much the bf16 rounding alone moves the gradient) ...
    exact = deepcopy(base).train()
    preds_exact = yolo_oracle.forward_with_grad(exact, x)
    g = torch.Generator().manual_seed(7)
    G = [torch.randn(p.shape, generator=g) / p.numel() ** 0.5 for p in preds_exact]
    sum((p * gg).sum() for p, gg in zip(preds_exact, G)).backward()
    # ... and the rounding-matched oracle the CUDA path is held to (bf16 at the CUDA path's storage points, fp32 backward)
    ref = deepcopy(base).train()
    yolo_oracle.SIMULATE_BF16 = True
    try:
        preds_ref = yolo_oracle.forward_with_grad(ref, x)
    finally:
        yolo_oracle.SIMULATE_BF16 = False
    sum((p * gg).sum() for p, gg in zip(preds_ref, G)).backward()
    cos0, ratio0, _ = _grad_stats(ref, exact)
    print(f"{name}: rounding-only noise floor (bf16-rounded vs exact fp32 oracle): gradient cosine {cos0:.4f}, norm ratio {ratio0:.4f}")
    m = deepcopy(base).cuda().train()
    preds = m(x.cuda())
    sum((p * gg.cuda()).sum() for p, gg in zip(preds, G)).backward()
    torch.cuda.synchronize()
    for a, b in zip(preds, preds_ref):
        rel = float((a.detach().cpu() - b.detach()).norm() / b.detach().norm())
        print(f"{name}: train-mode logits rel-L2 vs rounding-matched oracle {rel:.4f}")
        assert rel < (1e-2 if name.startswith("mini") else 4e-2)
    for (n1, b1), (n2, b2) in zip(m.named_buffers(), ref.named_buffers()):
        if n1.endswith("running_mean") or n1.endswith("running_var"):
            assert torch.allclose(b1.cpu(), b2, rtol=5e-2, atol=5e-3), n1
    cos, ratio, worst = _grad_stats(m, ref)
    print(f"{name}: fixed-upstream gradient cosine {cos:.5f}, norm ratio {ratio:.4f}, worst tensor {worst}")
    # criterion: the CUDA path agrees with the rounding-matched oracle at least as well as the bf16 rounding alone
    # moves the exact fp32 gradient (measured r1: 0.989 vs 0.976, 0.984 vs 0.971, 0.855 vs 0.723), plus an absolute floor
    assert cos > cos0 and cos > min_cos, f"global gradient cosine {cos} (noise floor {cos0}), worst tensor {worst}"
    assert 0.93 < ratio < 1.07, ratio


@pytest.mark.parametrize("name,hw,bs", [("yolov5n", (128, 128), 4), ("yolov5s", (160, 192), 2), ("yolov5_v5", (128, 128), 2)])
def test_train_step_matches_oracle(name, hw, bs):
    from ayolov2_b200 import synth
    from ayolov2_b200.loss import ComputeLoss
    from oracle import loss_oracle, yolo_oracle

    base = synth.build_model(name, seed=0)
    base.hyp = dict(HYP)
    x = torch.rand((bs, 3, *hw), generator=torch.Generator().manual_seed(3))
    targets = _targets(bs, 12, 4)
    # ---- oracle (CPU fp32, autograd)
    ref = deepcopy(base).train()
    preds_ref = yolo_oracle.forward_with_grad(ref, x)
    head = ref.model[-1]
    loss_ref, items_ref = loss_oracle.compute_loss(preds_ref, targets, head.anchors, HYP, head.nc)
    loss_ref.backward()
    # ---- CUDA
    m = deepcopy(base).cuda().train()
    preds = m(x.cuda())
    loss, items = ComputeLoss(m)(preds, targets.cuda())
    loss.backward()
    torch.cuda.synchronize()
    # Train-mode BatchNorm re-normalises every layer with batch statistics, which amplifies bf16 rounding: a CPU
    # simulation of this model with identical rounding points (bf16 weights/activations, fp32 math) differs from the
    # fp32 oracle by 2.4-5.2 % relative L2 on these inputs (max-normalised 8-17 %), so that is the noise floor.
    for a, b in zip(preds, preds_ref):
        rel = float((a.detach().cpu() - b.detach()).norm() / b.detach().norm())
        print(f"{name}: train-mode logits rel-L2 {rel:.4f}")
        assert rel < 8e-2, f"train-mode logits rel-L2 {rel}"
    print(f"{name}: loss {float(loss):.5f} vs oracle {float(loss_ref):.5f}; items {items.tolist()} vs {items_ref.tolist()}")
    assert abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)) < 3e-2, (float(loss), float(loss_ref))
    # BatchNorm running statistics were updated with batch statistics (momentum 0.03)
    for (n1, b1), (n2, b2) in zip(m.named_buffers(), ref.named_buffers()):
        if n1.endswith("running_mean") or n1.endswith("running_var"):
            assert torch.allclose(b1.cpu(), b2, rtol=5e-2, atol=5e-3), n1
    dots = n1s = n2s = 0.0
    worst = (0.0, "")
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        g1, g2 = p.grad.detach().cpu().double(), q.grad.detach().double()
        dots += float((g1 * g2).sum())
        n1s += float((g1 * g1).sum())
        n2s += float((g2 * g2).sum())
        rel = float((g1 - g2).norm() / (g2.norm() + 1e-12))
        if g2.norm() > 1e-6 and rel > worst[0]:
            worst = (rel, n)
    cos = dots / (n1s ** 0.5 * n2s ** 0.5)
    print(f"{name}: full-step gradient cosine {cos:.5f}, norm ratio {(n1s / n2s) ** 0.5:.4f}, worst tensor {worst}")
    # With the detection-prior head biases the loss gradient sigma(x) - t is exponentially sensitive to the logits
    # (a 4 % logit perturbation moves d(loss)/d(pred) by 25-50 %), so the end-to-end direction is only loosely pinned;
    # the exact backward is pinned by test_backward_matches_oracle_for_fixed_upstream_gradient and test_loss_gpu.py.
    assert cos > 0.3 and 0.7 < (n1s / n2s) ** 0.5 < 1.4


def test_full_size_backward_against_rounding_matched_oracle():
    """BASELINE configs[2]'s map sizes (640x640: 320^2 ... 20^2, every box shape / halo path the bs-128 step uses) at a batch
    the fp32 oracle can hold: the oracle runs ON THE GPU here (plain torch fp32 autograd, cuDNN) as the checker, with the
    bf16 rounding points of the CUDA path (yolo_oracle.SIMULATE_BF16), for a fixed upstream gradient."""
    from oracle import yolo_oracle

    bs = 16
    base = _build("yolov5s")
    x = torch.rand((bs, 3, 640, 640), generator=torch.Generator().manual_seed(5)).cuda()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # the oracle is fp32: cuDNN's default would run its convolutions in TF32
    try:
        _full_size_backward(base, x, bs)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


def _full_size_backward(base, x, bs):
    from oracle import yolo_oracle

    exact = deepcopy(base).cuda().train()
    preds_exact = yolo_oracle.forward_with_grad(exact, x)
    g = torch.Generator().manual_seed(7)
    G = [(torch.randn(p.shape, generator=g) / p.numel() ** 0.5).cuda() for p in preds_exact]
    sum((p * gg).sum() for p, gg in zip(preds_exact, G)).backward()
    del preds_exact
    ref = deepcopy(base).cuda().train()
    yolo_oracle.SIMULATE_BF16 = True
    try:
        preds_ref = yolo_oracle.forward_with_grad(ref, x)
    finally:
        yolo_oracle.SIMULATE_BF16 = False
    sum((p * gg).sum() for p, gg in zip(preds_ref, G)).backward()
    cos0, ratio0, _ = _grad_stats(ref, exact)
    del exact
    torch.cuda.empty_cache()
    m = deepcopy(base).cuda().train()
    preds = m(x)
    sum((p * gg).sum() for p, gg in zip(preds, G)).backward()
    torch.cuda.synchronize()
    for lv, (a, b) in enumerate(zip(preds, preds_ref)):
        rel = float((a.detach() - b.detach()).norm() / b.detach().norm())
        print(f"yolov5s 640x640 bs{bs}: train-mode logits P{3 + lv} rel-L2 vs rounding-matched oracle {rel:.4f}")
        assert rel < 4e-2, rel
    cos, ratio, worst = _grad_stats(m, ref)
    print(f"yolov5s 640x640 bs{bs}: fixed-upstream gradient cosine {cos:.5f} (rounding-only noise floor {cos0:.5f}), "
          f"norm ratio {ratio:.4f}, worst tensor {worst}")
    # measured (round 2): cosine 0.914 against a rounding-only noise floor of 0.736, norm ratio 1.001
    from _parity import record

    record(f"train_fullsize_yolov5s_640_b{bs}/fixed_upstream_gradient", cosine=cos, rounding_noise_floor_cosine=cos0, norm_ratio=ratio)
    assert cos > cos0 and cos > 0.85 and 0.93 < ratio < 1.07, (cos, cos0, ratio, worst)


def test_tucker_decomposed_model_trains():
    """Fine-tuning a Tucker-2 decomposed model (decompose_model.py writes such checkpoints; the reference trains them like
    any other): every kxk Conv is the nn.Sequential(1x1, kxk, 1x1) of decomposition.py:363-424. Backward for a fixed
    upstream gradient vs the fp32 oracle's autograd through the same modules."""
    from ayolov2_b200 import tucker
    from oracle import yolo_oracle

    base = _build("mini_v6")
    replaced = tucker.decompose_model_fixed(base, ratio=0.5)
    assert len(replaced) >= 3
    bs, hw = 4, (128, 160)
    x = torch.rand((bs, 3, *hw), generator=torch.Generator().manual_seed(3))
    ref = deepcopy(base).train()
    preds_ref = yolo_oracle.forward_with_grad(ref, x)
    g = torch.Generator().manual_seed(7)
    G = [torch.randn(p.shape, generator=g) / p.numel() ** 0.5 for p in preds_ref]
    sum((p * gg).sum() for p, gg in zip(preds_ref, G)).backward()
    m = deepcopy(base).cuda().train()
    preds = m(x.cuda())
    sum((p * gg.cuda()).sum() for p, gg in zip(preds, G)).backward()
    torch.cuda.synchronize()
    for a, b in zip(preds, preds_ref):
        rel = float((a.detach().cpu() - b.detach()).norm() / b.detach().norm())
        print(f"tucker mini_v6: train-mode logits rel-L2 vs fp32 oracle {rel:.4f}")
        assert rel < 8e-2, rel
    names = [n for n, _ in m.named_parameters()]
    assert any(".conv.0.weight" in n for n in names) and all(p.grad is not None for p in m.parameters())
    cos, ratio, worst = _grad_stats(m, ref)
    print(f"tucker mini_v6: fixed-upstream gradient cosine {cos:.5f} vs the exact fp32 oracle, norm ratio {ratio:.4f}, worst {worst}")
    assert cos > 0.95 and 0.93 < ratio < 1.07, (cos, ratio, worst)  # measured: 0.977, 1.005
