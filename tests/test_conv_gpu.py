"""Parity of the tcgen05 implicit-GEMM conv (ay2_conv_plan_*) on the GPU.

Checked against (a) a plain PyTorch fp32 conv2d of the same bf16-rounded operands and (b) the SIMT
reference kernel ay2_conv_reference_simt. Tolerance: the kernel accumulates bf16 products in fp32 and rounds
the result to bf16 once -> |err| <= 2^-8 * |ref| + small absolute slack for the accumulation order.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # (B, H, W, Cin, Cout, k, s, p, act, residual, in_slice, out_slice)
    (2, 16, 16, 64, 64, 1, 1, 0, 1, False, False, False),
    (2, 16, 16, 64, 64, 3, 1, 1, 1, False, False, False),
    (2, 32, 32, 16, 32, 3, 1, 1, 1, False, False, False),   # stem-like: CK=16 / SW32, N=32
    (2, 32, 32, 32, 64, 3, 2, 1, 1, False, False, False),   # CK=32 / SW64, stride 2
    (3, 20, 20, 128, 128, 3, 1, 1, 1, True, False, False),  # 4x4 boxes, residual
    (2, 40, 40, 64, 64, 3, 1, 1, 1, True, True, True),      # 8x8 boxes, slices of wider buffers, in-place style
    (2, 40, 40, 128, 256, 3, 2, 1, 1, False, False, True),  # stride 2 -> 20x20
    (1, 20, 20, 512, 512, 1, 1, 0, 1, False, True, False),  # two N tiles
    (2, 20, 20, 256, 256, 1, 1, 0, 0, False, False, False), # head-like (255 padded to 256), no act
    (1, 24, 40, 64, 128, 3, 1, 1, 0, False, False, False),  # ragged boxes (24x40)
    (5, 10, 10, 256, 128, 3, 1, 1, 1, False, False, False), # tiny maps, M tail
    (1, 80, 80, 128, 128, 3, 2, 1, 1, False, True, True),
    (2, 20, 20, 1024, 512, 1, 1, 0, 1, False, False, False),# SPPF conv2-like, long K
    # 3x3 / s1 / p1 with Cin % 64 == 0 -> halo kernel (16 x 8 half tiles, shifted-descriptor taps)
    (2, 40, 40, 128, 128, 3, 1, 1, 1, True, False, False),  # two K chunks, residual, ragged bands (40 = 16 + 16 + 8)
    (1, 80, 80, 64, 64, 3, 1, 1, 1, False, True, True),     # exact tiling, slices of wider buffers
    (1, 48, 24, 64, 32, 3, 1, 1, 0, False, False, False),   # odd number of halves (2 + 1), N = 32
    (1, 32, 32, 128, 256, 3, 1, 1, 1, False, False, False), # two N tiles of 128
    (2, 16, 40, 320, 64, 3, 1, 1, 1, True, False, True),    # five K chunks
]


# CTA-pair form (tcgen05 cta_group::2): N tile 128 / 256, K >= 256, at least one M tile per SM (>= 148 x 128 output pixels)
PAIR_CASES = [
    (2, 112, 112, 256, 256, 1, 1, 0, 1, False, False, False),   # 1x1, 196 M tiles
    (2, 112, 96, 256, 128, 1, 1, 0, 1, True, True, True),       # N tile 128, residual, channel slices, 168 M tiles
    (3, 160, 160, 64, 128, 3, 2, 1, 1, False, False, False),    # 3x3 stride 2, K = 576, 150 M tiles
    (1, 160, 160, 128, 256, 3, 2, 1, 1, False, False, True),    # 3x3 stride 2, 50 M tiles -> too few: single-CTA form
    (1, 140, 140, 512, 512, 1, 1, 0, 0, False, False, False),   # two N tiles, ODD number of M tiles (154 boxes of 8x16 -> 153.1)
    (7, 53, 61, 256, 248, 1, 1, 0, 0, False, False, False),     # 248 of 256 tile columns, ragged boxes, batch straddling pairs
]


def _run_case(case, seed=0):
    from ayolov2_b200 import ops

    B, H, W, Cin, Cout, k, s, p, act, use_res, in_slice, out_slice = case
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    cs_in = Cin + 32 if in_slice else Cin
    c0_in = 16 if in_slice else 0
    cout8 = (Cout + 7) // 8 * 8
    cs_out = cout8 + 64 if out_slice else cout8
    c0_out = 24 if out_slice else 0
    xbuf = torch.randn((B, H, W, cs_in), device=dev, generator=g).to(torch.bfloat16)
    ybuf = torch.full((B, OH, OW, cs_out), 7.0, device=dev, dtype=torch.bfloat16)
    x = ops.ActView(xbuf, c0_in, Cin)
    y = ops.ActView(ybuf, c0_out, Cout)
    w = torch.randn((Cout, Cin, k, k), device=dev, generator=g) * (1.0 / (Cin * k * k) ** 0.5)
    b = torch.randn(Cout, device=dev, generator=g) * 0.5
    wp, bp = ops.pack_conv_weight(w, b)
    res = None
    if use_res:
        rbuf = torch.randn((B, OH, OW, cs_out), device=dev, generator=g).to(torch.bfloat16)
        res = ops.ActView(rbuf, c0_out, Cout)
    plan = ops.ConvPlan(x, y, wp, bp, k, k, s, p, act, residual=res)
    plan.run()
    torch.cuda.synchronize()
    got = y.tensor().float().clone()
    if seed == 1:  # a second run of the same plan (persistent barriers / TMEM are re-initialised per launch)
        plan.run()
        torch.cuda.synchronize()
        assert torch.equal(got, y.tensor().float())
    untouched = ybuf.clone()
    untouched[..., c0_out:c0_out + Cout] = 7.0
    assert torch.all(untouched == 7.0), "conv wrote outside its channel slice"

    # (a) torch fp32 reference on the same rounded operands
    xin = x.tensor().float().permute(0, 3, 1, 2)
    wr = wp[:Cout].float().view(Cout, k, k, Cin).permute(0, 3, 1, 2)
    ref = F.conv2d(xin, wr, bp[:Cout], stride=s, padding=p)
    if act == 1:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    if use_res:
        ref = ref + res.tensor().float()
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2
    bad = err > tol
    assert not bad.any(), (f"case {case}: {int(bad.sum())} / {bad.numel()} mismatches, max err {float(err.max()):.4f}, "
                           f"first bad index {bad.nonzero()[0].tolist()}")

    # (b) SIMT reference kernel (same rounding points) -> at most 1 bf16 ulp apart
    ybuf.fill_(7.0)
    plan.run_reference_simt()
    torch.cuda.synchronize()
    simt = y.tensor().float()
    err2 = (got - simt).abs()
    assert float((err2 - (2.0 ** -7 * simt.abs() + 1e-2)).max()) <= 0, f"case {case}: SIMT mismatch {float(err2.max())}"
    return plan


@pytest.mark.parametrize("case", CASES, ids=[f"c{i}" for i in range(len(CASES))])
def test_conv_parity(case):
    _run_case(case)


@pytest.mark.parametrize("case", PAIR_CASES, ids=[f"p{i}" for i in range(len(PAIR_CASES))])
def test_conv_parity_cta_pair(case):
    """The same parity bar for the cta_group::2 form; the plan reports which form it chose."""
    plan = _run_case(case, seed=1)
    out_px = case[0] * ((case[1] + 2 * case[7] - case[5]) // case[6] + 1) * ((case[2] + 2 * case[7] - case[5]) // case[6] + 1)
    n_tile, spatial = case[4], case[5] > 1
    if out_px >= 150 * 128 and (n_tile > 128 or spatial):  # (N tiles of 128 pair up only under a spatial kernel)
        assert plan.pair, "expected the CTA-pair form"


def test_conv_inplace_residual():
    """Bottleneck shortcut: y1 <- y1 + conv3x3(t), residual and output are the same slice."""
    from ayolov2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    B, H, W, C_ = 2, 40, 40, 64
    buf = torch.randn((B, H, W, 2 * C_), device="cuda", generator=g).to(torch.bfloat16)
    t = torch.randn((B, H, W, C_), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((C_, C_, 3, 3), device="cuda", generator=g) * 0.05
    wp, bp = ops.pack_conv_weight(w, None)
    y1 = ops.ActView(buf, 0, C_)
    before = buf.clone()
    plan = ops.ConvPlan(ops.ActView(t, 0, C_), y1, wp, bp, 3, 3, 1, 1, 1, residual=y1)
    plan.run()
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(t.float().permute(0, 3, 1, 2), wp.float().view(C_, 3, 3, C_).permute(0, 3, 1, 2), bp, padding=1))
    ref = ref.permute(0, 2, 3, 1) + before[..., :C_].float()
    err = (buf[..., :C_].float() - ref).abs()
    assert float((err - (2.0 ** -7 * ref.abs() + 2e-2)).max()) <= 0
    assert torch.equal(buf[..., C_:], before[..., C_:])


@pytest.mark.parametrize("B,H,W,Cout", [(2, 32, 48, 32), (2, 24, 40, 32), (1, 20, 20, 64), (3, 40, 72, 64)])
def test_conv_window_mode_packed_stem(B, H, W, Cout):
    """16-channel 3x3 conv run as 3x1 taps over overlapping 4-pixel windows (in_pix_stride 16 < cin 64) of a
    horizontally padded buffer == plain 3x3 conv of the 16-channel tensor. Since round 2 this form runs in the halo kernel
    (one 18-line tile per item, the three vertical taps are shifted descriptors): full tiles, a one-half right edge (W = 40,
    72), partial bands (H = 24, 40) and a map the halo tiles fill badly (20 x 20: generic kernel)."""
    from ayolov2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(4)
    Wp = W + 8
    buf = torch.zeros((B, H, Wp, 16), device="cuda", dtype=torch.bfloat16)
    x = torch.randn((B, H, W, 16), device="cuda", generator=g).to(torch.bfloat16)
    buf[:, :, 1:W + 1] = x
    w = torch.randn((Cout, 16, 3, 3), device="cuda", generator=g) * 0.1
    ww = torch.zeros((Cout, 64, 3, 1), device="cuda")
    for kw in range(3):
        ww[:, kw * 16:(kw + 1) * 16, :, 0] = w[:, :, :, kw]
    wp, bp = ops.pack_conv_weight(ww, None)
    y = ops.new_act(B, H, W, Cout)
    plan = ops.ConvPlan(ops.ActView(buf, 0, 16), y, wp, bp, 3, 1, 1, 1, 1, pad_w=0, window=(64, W, 16, Wp))
    plan.run()
    torch.cuda.synchronize()
    got = y.tensor().float().clone()
    wr = w.to(torch.bfloat16).float()
    ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), wr, None, padding=1)).permute(0, 2, 3, 1)
    err = (got - ref).abs()
    assert float((err - (2.0 ** -7 * ref.abs() + 2e-2)).max()) <= 0, float(err.max())
    plan.run_reference_simt()
    torch.cuda.synchronize()
    assert float((y.tensor().float() - got).abs().max()) < 3e-2


@pytest.mark.parametrize("B,H,W,Cout", [(2, 32, 48, 64), (1, 40, 72, 64), (3, 16, 24, 128)])
def test_conv_pixel_pair_form_stride2(B, H, W, Cout):
    """3x3 / stride 2 / pad 1 over 32 channels run as a 3x2-tap conv over PAIRS of pixels (the NHWC buffer viewed as
    [B, H, W/2, 64]; stride 2 over rows, 1 over pairs; `stride_w` in ay2_conv_desc) == the plain convolution."""
    from ayolov2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn((B, H, W, 32), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((Cout, 32, 3, 3), device="cuda", generator=g) * 0.1
    bias = torch.randn(Cout, device="cuda", generator=g) * 0.1
    wp, bp = ops.pack_conv_weight(ops.pixel_pair_weight(w), bias)
    y = ops.new_act(B, H // 2, W // 2, Cout)
    xp = ops.ActView(x.view(B, H, W // 2, 64), 0, 64)
    plan = ops.ConvPlan(xp, y, wp, bp, 3, 2, 2, 1, 1, pad_w=1, stride_w=1)
    plan.run()
    torch.cuda.synchronize()
    got = y.tensor().float()
    ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), bias, stride=2, padding=1)).permute(0, 2, 3, 1)
    err = (got - ref).abs()
    assert float((err - (2.0 ** -7 * ref.abs() + 2e-2)).max()) <= 0, float(err.max())
