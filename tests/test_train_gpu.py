"""GPU parity of the training-step kernels (BatchNorm batch statistics + SiLU forward/backward, conv weight
gradient on tcgen05, data-movement backward, SGD+EMA) against plain PyTorch fp32 autograd on the same
bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand_act(B, H, W, C, cs=None, c0=0, seed=0, scale=1.0):
    from ayolov2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(seed)
    cs = cs or C
    buf = (torch.randn((B, H, W, cs), device="cuda", generator=g) * scale).to(torch.bfloat16)
    return ops.ActView(buf, c0, C)


@pytest.mark.parametrize("C,cs,c0", [(64, 64, 0), (32, 96, 32), (256, 256, 0)])
def test_bn_silu_forward_backward(C, cs, c0):
    from ayolov2_b200 import ops

    B, H, W = 4, 20, 24
    z = _rand_act(B, H, W, C, cs, c0, seed=1, scale=2.0)
    dy = _rand_act(B, H, W, C, seed=2)
    g = torch.Generator(device="cuda").manual_seed(3)
    gamma = 0.5 + torch.rand(C, device="cuda", generator=g)
    beta = 0.3 * torch.randn(C, device="cuda", generator=g)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    mean, invstd = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    scratch = torch.empty(2 * C, dtype=torch.float64, device="cuda")
    ops.bn_batch_stats(z, 1e-3, 0.03, rm, rv, scratch, mean, invstd)
    y = ops.new_act(B, H, W, C)
    ops.bn_act_fwd(z, mean, invstd, gamma, beta, ops.ACT_SILU, y)
    dz = ops.new_act(B, H, W, C)
    ops.bn_act_bwd(dy, z, mean, invstd, gamma, beta, ops.ACT_SILU, scratch, dz)
    torch.cuda.synchronize()
    # reference
    zt = z.tensor().float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gt, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    yt = F.silu(F.batch_norm(zt, rm2, rv2, gt, bt, training=True, momentum=0.03, eps=1e-3))
    yt.backward(dy.tensor().float().permute(0, 3, 1, 2))
    assert torch.allclose(rm, rm2, atol=1e-5) and torch.allclose(rv, rv2, rtol=1e-4, atol=1e-5)
    assert torch.allclose(y.tensor().float(), yt.detach().permute(0, 2, 3, 1), rtol=2e-2, atol=2e-2)
    ref_dz = zt.grad.permute(0, 2, 3, 1)
    err = (dz.tensor().float() - ref_dz).abs().max() / ref_dz.abs().max()
    assert float(err) < 2e-2, float(err)
    assert torch.allclose(scratch[:C].float(), bt.grad, rtol=1e-3, atol=1e-3)
    assert torch.allclose(scratch[C:].float(), gt.grad, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("case", [
    # B, H, W, Cin, Cout, k, s, p, in_slice
    (2, 16, 16, 64, 64, 1, 1, 0, False),
    (2, 16, 16, 64, 128, 3, 1, 1, False),
    (3, 20, 20, 128, 64, 3, 1, 1, True),
    (2, 32, 32, 32, 64, 3, 2, 1, False),
    (2, 40, 40, 256, 128, 1, 1, 0, True),
    (4, 20, 20, 512, 256, 1, 1, 0, False),
    (2, 16, 24, 32, 32, 3, 1, 1, False),
    (2, 40, 40, 64, 128, 3, 2, 1, False),
    # 3x3 / s1: the halo-tile form of the kernel (one (BH+2) x (BW+2) x tile per box, taps = shifted descriptors) with
    # 8 x 16 boxes (160, 80), two 8 x 8 boxes per chunk (40), four 4 x 8 (20, above), ragged maps and 16 input channels
    (1, 160, 160, 32, 32, 3, 1, 1, False),
    (2, 80, 80, 64, 64, 3, 1, 1, True),
    (2, 40, 40, 128, 128, 3, 1, 1, False),
    (2, 12, 20, 16, 32, 3, 1, 1, False),
    (3, 22, 38, 32, 48, 3, 1, 1, False),
], ids=lambda c: "x".join(map(str, c)))
def test_conv_wgrad(case):
    from ayolov2_b200 import ops

    B, H, W, Cin, Cout, k, s, p, in_slice = case
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = _rand_act(B, H, W, Cin, Cin + 64 if in_slice else Cin, 32 if in_slice else 0, seed=5)
    dz = _rand_act(B, OH, OW, Cout, seed=6)
    dw = torch.zeros((Cout, k * k * Cin), device="cuda")
    ops.conv_wgrad(x, dz, dw, k, k, s, p)
    ops.conv_wgrad(x, dz, dw, k, k, s, p)  # accumulates: expect 2x
    torch.cuda.synchronize()
    xt = x.tensor().float().permute(0, 3, 1, 2)
    w = torch.zeros((Cout, Cin, k, k), device="cuda", requires_grad=True)
    F.conv2d(xt, w, None, stride=s, padding=p).backward(dz.tensor().float().permute(0, 3, 1, 2))
    ref = 2.0 * w.grad.permute(0, 2, 3, 1).reshape(Cout, -1)
    err = (dw - ref).abs().max() / ref.abs().max()
    assert float(err) < 2e-3, f"{case}: {float(err)}"


def test_upsample_and_maxpool_backward():
    from ayolov2_b200 import ops

    B, H, W, C = 2, 10, 12, 32
    dy = _rand_act(B, 2 * H, 2 * W, C, seed=7)
    dx = ops.new_act(B, H, W, C)
    dx.buf.fill_(1.0)
    ops.upsample2x_bwd(dy, dx, accumulate=True)
    torch.cuda.synchronize()
    xt = torch.zeros((B, C, H, W), device="cuda", requires_grad=True)
    F.interpolate(xt, scale_factor=2.0, mode="nearest").backward(dy.tensor().float().permute(0, 3, 1, 2))
    assert torch.allclose(dx.tensor().float(), 1.0 + xt.grad.permute(0, 2, 3, 1), rtol=2e-2, atol=2e-2)
    # max pool (bf16 inputs have many ties: the first maximum in window order must receive the gradient)
    x = _rand_act(B, 13, 11, C, seed=8)
    g = _rand_act(B, 13, 11, C, seed=9)
    d = ops.new_act(B, 13, 11, C)
    ops.maxpool_bwd(x, g, 5, d, accumulate=False)
    torch.cuda.synchronize()
    xt = x.tensor().float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.max_pool2d(xt, 5, 1, 2).backward(g.tensor().float().permute(0, 3, 1, 2))
    assert torch.allclose(d.tensor().float(), xt.grad.permute(0, 2, 3, 1), rtol=2e-2, atol=3e-2)


def test_head_layout_roundtrip_and_channel_sum():
    from ayolov2_b200 import ops

    B, na, ny, nx, no = 2, 3, 5, 7, 85
    g = torch.randn((B, na, ny, nx, no), device="cuda")
    out = ops.new_act(B, ny, nx, 256)
    ops.head_grad_to_nhwc(g, out)
    back = torch.empty_like(g)
    ops.head_logits_to_train(out, na, no, back)
    sums = torch.zeros(256, dtype=torch.float64, device="cuda")
    ops.channel_sum(out, sums)
    torch.cuda.synchronize()
    assert torch.equal(back, g.to(torch.bfloat16).float())
    assert torch.all(out.buf[..., 255] == 0)
    ref = out.buf.float().sum((0, 1, 2)).double()
    assert torch.allclose(sums, ref, rtol=1e-5, atol=1e-4)


def test_sgd_nesterov_ema_matches_torch():
    from ayolov2_b200 import ops

    torch.manual_seed(0)
    p = torch.randn(10000, device="cuda")
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.SGD([ref], lr=0.01, momentum=0.937, nesterov=True, weight_decay=5e-4)
    mom = torch.zeros_like(p)
    ema = p.clone()
    ema_ref = p.clone()
    for step in range(3):
        g = torch.randn_like(p)
        ref.grad = g.clone()
        opt.step()
        ema_ref = 0.99 * ema_ref + 0.01 * ref.detach()
        ops.sgd_ema_step(p, g, mom, ema, 0.01, 0.937, 5e-4, True, 0.99)
    torch.cuda.synchronize()
    assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(ema, ema_ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", [(2, 16, 16, 64, 64, 3, 1, 1), (2, 20, 20, 128, 64, 1, 1, 0), (2, 32, 32, 32, 64, 3, 2, 1),
                                  (2, 40, 40, 64, 128, 3, 2, 1)], ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("accumulate", [False, True])
def test_conv_dgrad_via_forward_kernel(case, accumulate):
    from ayolov2_b200 import ops

    B, H, W, Cin, Cout, k, s, p = case
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    g = torch.Generator(device="cuda").manual_seed(11)
    w = torch.randn((Cout, Cin, k, k), device="cuda", generator=g) * 0.1
    dz = _rand_act(B, OH, OW, Cout, seed=12)
    dx = _rand_act(B, H, W, Cin, seed=13)
    before = dx.tensor().float().clone()
    for plan in ops.make_dgrad_plans(dz, dx, w, s, p, accumulate):
        plan.run()
    torch.cuda.synchronize()
    xt = torch.zeros((B, Cin, H, W), device="cuda", requires_grad=True)
    F.conv2d(xt, w.to(torch.bfloat16).float(), None, stride=s, padding=p).backward(dz.tensor().float().permute(0, 3, 1, 2))
    ref = xt.grad.permute(0, 2, 3, 1) + (before if accumulate else 0.0)
    err = (dx.tensor().float() - ref).abs().max() / ref.abs().max()
    assert float(err) < 2e-2, float(err)
