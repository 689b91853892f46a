"""GPU parity of the fused CUDA loss (ay2_yolo_loss) with the CPU oracle (oracle/loss_oracle.py, pinned to the
reference ComputeLoss) and with the committed golden fixture: loss, items and d(loss)/d(preds).
Tolerance: fp32 1e-3 relative (BASELINE.json north_star); observed ~1e-5."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "loss_golden.npz")
HYP_KEYS = sorted(["box", "cls", "cls_pw", "obj", "obj_pw", "anchor_t", "fl_gamma", "label_smoothing"])
ANCHORS = torch.tensor([[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]).float().view(3, 3, 2)


class _Head:
    def __init__(self, nc):
        self.nl, self.na, self.nc = 3, 3, nc
        self.stride = torch.tensor([8.0, 16.0, 32.0])
        self.anchors = ANCHORS / self.stride.view(-1, 1, 1)


class _Model(torch.nn.Module):
    def __init__(self, nc, hyp):
        super().__init__()
        self.hyp = dict(hyp)
        self.model = [_Head(nc)]


def _check(preds_cpu, targets, hyp, nc, scale=1.0):
    from ayolov2_b200.loss import ComputeLoss
    from oracle import loss_oracle

    head = _Head(nc)
    p_ref = [p.clone().requires_grad_(True) for p in preds_cpu]
    l_ref, it_ref = loss_oracle.compute_loss(p_ref, targets, head.anchors, hyp, nc)
    (l_ref * scale).backward()
    p_gpu = [p.clone().cuda().requires_grad_(True) for p in preds_cpu]
    fn = ComputeLoss(_Model(nc, hyp))
    l_gpu, it_gpu = fn(p_gpu, targets.cuda())
    (l_gpu * scale).backward()
    torch.cuda.synchronize()
    assert torch.allclose(l_gpu.cpu(), l_ref.detach(), rtol=1e-4), (l_gpu, l_ref)
    assert torch.allclose(it_gpu.cpu(), it_ref, rtol=1e-4, atol=1e-6), (it_gpu, it_ref)
    for a, b in zip(p_gpu, p_ref):
        ga, gb = a.grad.cpu(), b.grad
        denom = gb.abs().max().clamp_min(1e-12)
        assert float((ga - gb).abs().max() / denom) < 1e-3, float((ga - gb).abs().max() / denom)
    return l_gpu, it_gpu


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_loss_golden(case):
    z = np.load(GOLD)
    hyp = dict(zip(HYP_KEYS, z[f"c{case}_hyp"].tolist()))
    preds = [torch.from_numpy(z[f"c{case}_pred{i}"]) for i in range(3)]
    targets = torch.from_numpy(z[f"c{case}_targets"])
    nc = preds[0].shape[-1] - 5
    l, it = _check(preds, targets, hyp, nc)
    assert torch.allclose(l.cpu(), torch.from_numpy(z[f"c{case}_loss"]), rtol=1e-4)
    assert torch.allclose(it.cpu(), torch.from_numpy(z[f"c{case}_items"]), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("seed,bs,nt,scale", [(0, 8, 60, 1.0), (1, 4, 200, 8.0), (2, 2, 1, 1.0)])
def test_loss_random(seed, bs, nt, scale):
    g = torch.Generator().manual_seed(seed)
    nc = 80
    preds = [torch.randn(bs, 3, 160 // s, 160 // s, nc + 5, generator=g) for s in (8, 16, 32)]
    t = torch.zeros(nt, 6)
    t[:, 0] = torch.randint(0, bs, (nt,), generator=g).float()
    t[:, 1] = torch.randint(0, nc, (nt,), generator=g).float()
    t[:, 2:4] = 0.02 + 0.96 * torch.rand(nt, 2, generator=g)
    t[:, 4:6] = torch.exp(np.log(0.02) + (np.log(0.8) - np.log(0.02)) * torch.rand(nt, 2, generator=g))
    hyp = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, label_smoothing=0.0)
    _check(preds, t, hyp, nc, scale)


@pytest.mark.parametrize("gamma", [1.5, 2.0])
def test_loss_focal(gamma):
    """hyp["fl_gamma"] > 0: both BCE criteria become FocalLoss(BCE, gamma) (losses.py:193-196); forward and gradient vs the
    pinned oracle, same tolerances as the plain path."""
    g = torch.Generator().manual_seed(11)
    nc, bs, nt = 80, 4, 50
    preds = [torch.randn(bs, 3, 160 // s, 160 // s, nc + 5, generator=g) * 2 for s in (8, 16, 32)]
    t = torch.zeros(nt, 6)
    t[:, 0] = torch.randint(0, bs, (nt,), generator=g).float()
    t[:, 1] = torch.randint(0, nc, (nt,), generator=g).float()
    t[:, 2:4] = 0.05 + 0.9 * torch.rand(nt, 2, generator=g)
    t[:, 4:6] = torch.exp(np.log(0.03) + (np.log(0.7) - np.log(0.03)) * torch.rand(nt, 2, generator=g))
    hyp = dict(box=0.05, cls=0.5, cls_pw=0.9, obj=1.0, obj_pw=1.1, anchor_t=4.0, fl_gamma=gamma, label_smoothing=0.05)
    _check(preds, t, hyp, nc)


def test_loss_autobalance():
    """autobalance=True (losses.py:209,286-292): the level weights evolve like the oracle's over three calls, every loss
    and the gradient of the LAST call (taken after its weights were already updated) match the oracle."""
    from ayolov2_b200.loss import ComputeLoss
    from oracle import loss_oracle

    nc, bs, nt = 80, 2, 12
    hyp = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, label_smoothing=0.0)
    head = _Head(nc)
    fn = ComputeLoss(_Model(nc, hyp), autobalance=True)
    assert fn.ssi == 1
    bal = [4.0, 1.0, 0.4]
    for seed in (7, 8, 9):
        g = torch.Generator().manual_seed(seed)
        preds = [torch.randn(bs, 3, 160 // s, 160 // s, nc + 5, generator=g) for s in (8, 16, 32)]
        t = torch.zeros(nt, 6)
        t[:, 0] = torch.randint(0, bs, (nt,), generator=g).float()
        t[:, 1] = torch.randint(0, nc, (nt,), generator=g).float()
        t[:, 2:4] = 0.05 + 0.9 * torch.rand(nt, 2, generator=g)
        t[:, 4:6] = torch.exp(np.log(0.03) + (np.log(0.7) - np.log(0.03)) * torch.rand(nt, 2, generator=g))
        p_ref = [p.clone().requires_grad_(True) for p in preds]
        l_ref, it_ref = loss_oracle.compute_loss(p_ref, t, head.anchors, hyp, nc, balance=bal, ssi=1)
        p_gpu = [p.clone().cuda().requires_grad_(True) for p in preds]
        l_gpu, it_gpu = fn(p_gpu, t.cuda())
        assert torch.allclose(l_gpu.cpu(), l_ref.detach(), rtol=1e-4) and torch.allclose(it_gpu.cpu(), it_ref, rtol=1e-4, atol=1e-6)
        assert np.allclose(fn.balance, bal, rtol=1e-5)
    l_ref.backward()
    l_gpu.backward()
    for a, b in zip(p_gpu, p_ref):
        denom = b.grad.abs().max().clamp_min(1e-12)
        assert float((a.grad.cpu() - b.grad).abs().max() / denom) < 1e-3


def test_loss_duplicate_cells_last_wins():
    """Two identical targets hit the same cells: tobj takes the later candidate's IoU, gradients accumulate."""
    g = torch.Generator().manual_seed(3)
    nc = 4
    preds = [torch.randn(1, 3, 64 // s, 64 // s, nc + 5, generator=g) for s in (8, 16, 32)]
    t = torch.tensor([[0, 1, 0.52, 0.48, 0.3, 0.25], [0, 2, 0.52, 0.48, 0.3, 0.25], [0, 1, 0.5, 0.5, 0.1, 0.1]])
    hyp = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, label_smoothing=0.0)
    _check(preds, t, hyp, nc)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_loss_half_precision_preds(dtype):
    """The reference trainer calls ComputeLoss under amp.autocast with a torch head (yolo_trainer.py:318-324): preds arrive in
    fp16 / bf16. Forward uses the fp32 value of the half logits; the gradient comes back in the predictions' own dtype and
    equals the fp32 gradient at those logits to half rounding (no out-of-bounds write into half-sized buffers)."""
    from ayolov2_b200.loss import ComputeLoss

    g = torch.Generator().manual_seed(21)
    nc, bs, nt = 80, 3, 40
    preds = [torch.randn(bs, 3, 160 // s, 160 // s, nc + 5, generator=g).to(dtype) for s in (8, 16, 32)]
    t = torch.zeros(nt, 6)
    t[:, 0] = torch.randint(0, bs, (nt,), generator=g).float()
    t[:, 1] = torch.randint(0, nc, (nt,), generator=g).float()
    t[:, 2:4] = 0.05 + 0.9 * torch.rand(nt, 2, generator=g)
    t[:, 4:6] = torch.exp(np.log(0.03) + (np.log(0.7) - np.log(0.03)) * torch.rand(nt, 2, generator=g))
    hyp = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, label_smoothing=0.0)
    fn = ComputeLoss(_Model(nc, hyp))
    p_half = [p.clone().cuda().requires_grad_(True) for p in preds]
    canary = torch.full((4096,), 7.0, device="cuda")  # allocated right after the predictions: a stray fp32-sized write lands here
    l_half, it_half = fn(p_half, t.cuda())
    l_half.backward()
    p_f32 = [p.float().cuda().requires_grad_(True) for p in preds]
    l_f32, it_f32 = fn(p_f32, t.cuda())
    l_f32.backward()
    torch.cuda.synchronize()
    assert torch.equal(l_half, l_f32) and torch.equal(it_half, it_f32)
    assert bool((canary == 7.0).all())
    for a, b in zip(p_half, p_f32):
        assert a.grad.dtype == dtype and a.grad.shape == a.shape
        assert torch.equal(a.grad, b.grad.to(dtype))


def test_loss_noncontiguous_preds():
    """Predictions that are permuted views (a torch head's x.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2) without
    .contiguous()): the gradient must come back laid out for the view, identical to the contiguous case."""
    from ayolov2_b200.loss import ComputeLoss

    g = torch.Generator().manual_seed(22)
    nc, bs = 6, 2
    base = [torch.randn(bs, 3, nc + 5, 64 // s, 64 // s, generator=g).cuda().requires_grad_(True) for s in (8, 16, 32)]
    t = torch.tensor([[0, 1, 0.52, 0.48, 0.3, 0.25], [1, 2, 0.3, 0.6, 0.2, 0.4]])
    hyp = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, label_smoothing=0.0)
    fn = ComputeLoss(_Model(nc, hyp))
    l1, _ = fn([b.permute(0, 1, 3, 4, 2) for b in base], t.cuda())
    l1.backward()
    contig = [b.detach().permute(0, 1, 3, 4, 2).contiguous().requires_grad_(True) for b in base]
    l2, _ = fn(contig, t.cuda())
    l2.backward()
    assert torch.equal(l1, l2)
    for b, c in zip(base, contig):
        assert torch.equal(b.grad.permute(0, 1, 3, 4, 2), c.grad)
