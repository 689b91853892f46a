"""Parity of the fused convolution chain (ay2_chain_plan_*: 1x1 -> 3x3 [-> 1x1] in one kernel) on the GPU.

Reference = plain PyTorch fp32 convolutions of the same bf16-rounded weights, with the intermediates rounded to bf16
at the points where the kernel stores them (shared memory T and U) -- the same rounding points as the three-launch
path, whose intermediates are bf16 tensors in HBM. Tolerance: one final bf16 rounding + fp32 accumulation-order slack.
Covers the two users: Tucker-2 chains (decomposition.py:363-424: no bias / activation on the first two links) and the
kindle Bottleneck (SiLU on both links, shortcut add).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # (B, H, W, Cin, C1, C2, C3, acts, residual, in_slice, out_slice)
    (2, 16, 8, 64, 64, 64, 0, (1, 1, 0), True, False, False),     # one tile per image, Bottleneck
    (2, 32, 32, 32, 32, 32, 0, (1, 1, 0), True, True, True),      # CK = 32 (SW64), slices of wider buffers
    (1, 40, 40, 128, 128, 128, 0, (1, 1, 0), True, False, False), # ragged rows (40 = 2*16 + 8), 1 CTA/SM config
    (2, 24, 20, 64, 64, 64, 0, (1, 1, 0), False, False, False),   # ragged rows and columns, no shortcut
    (2, 32, 16, 64, 32, 48, 128, (0, 0, 1), False, False, False), # Tucker: ranks 32 / 48, SiLU on the last link
    (1, 48, 40, 128, 64, 64, 256, (0, 0, 1), False, True, False), # Tucker, N3 = 256
    (2, 16, 16, 256, 96, 80, 320, (0, 0, 1), False, False, False),# Tucker, two N3 tiles of 160, CK2 = 32, CK3 = 16
    (1, 20, 20, 64, 16, 16, 64, (0, 0, 0), True, False, True),    # tiny ranks, residual after a 3-link chain
    (3, 80, 80, 64, 64, 64, 0, (1, 1, 0), True, True, True),      # many tiles per CTA (persistence, ring wrap-around)
]


def _bf16(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("case", CASES, ids=[f"c{i}" for i in range(len(CASES))])
def test_chain_parity(case):
    from ayolov2_b200 import ops

    B, H, W, Cin, C1, C2, C3, acts, use_res, in_slice, out_slice = case
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1234)
    cout = C3 or C2
    cs_in, c0_in = (Cin + 32, 16) if in_slice else (Cin, 0)
    cs_out, c0_out = (cout + 64, 24) if out_slice else (cout, 0)
    xbuf = torch.randn((B, H, W, cs_in), device=dev, generator=g).to(torch.bfloat16)
    ybuf = torch.full((B, H, W, cs_out), 7.0, device=dev, dtype=torch.bfloat16)
    x, y = ops.ActView(xbuf, c0_in, Cin), ops.ActView(ybuf, c0_out, cout)
    w1 = torch.randn((C1, Cin, 1, 1), device=dev, generator=g) / Cin ** 0.5
    w2 = torch.randn((C2, C1, 3, 3), device=dev, generator=g) / (9 * C1) ** 0.5
    b1 = torch.randn(C1, device=dev, generator=g) * 0.5 if acts[0] else None  # Tucker links 1, 2 carry no bias
    b2 = torch.randn(C2, device=dev, generator=g) * 0.5 if (acts[1] or not C3) else None
    links = [ops.pack_chain_weight(w1, b1), ops.pack_chain_weight(w2, b2)]
    if C3:
        w3 = torch.randn((C3, C2, 1, 1), device=dev, generator=g) / C2 ** 0.5
        b3 = torch.randn(C3, device=dev, generator=g) * 0.5
        links.append(ops.pack_chain_weight(w3, b3))
    res = None
    if use_res:
        if Cin == cout and not out_slice and not in_slice:
            res = x  # the Bottleneck shortcut: residual == input
        else:
            rbuf = torch.randn((B, H, W, cs_out), device=dev, generator=g).to(torch.bfloat16)
            res = ops.ActView(rbuf, c0_out, cout)
    assert ops.chain_supported(ops.chain_desc(x, C1, C2, C3, *acts, y.cstride, res.cstride if res is not None else 0))
    plan = ops.ChainPlan(x, y, links, acts[:len(links)], residual=res)
    plan.run()
    plan.run()  # idempotent (persistent barriers / phases restart cleanly per launch)
    torch.cuda.synchronize()
    got = y.tensor().float().clone()
    untouched = ybuf.clone()
    untouched[..., c0_out:c0_out + cout] = 7.0
    assert torch.all(untouched == 7.0), "chain wrote outside its channel slice"

    def act(t, a):
        return F.silu(t) if a else t

    xin = x.tensor().float().permute(0, 3, 1, 2)
    t = _bf16(act(F.conv2d(xin, links[0][0].float().view(C1, 1, 1, Cin).permute(0, 3, 1, 2), links[0][1]), acts[0]))
    u = act(F.conv2d(t, links[1][0].float().view(C2, 3, 3, C1).permute(0, 3, 1, 2), links[1][1], padding=1), acts[1])
    if C3:
        u = act(F.conv2d(_bf16(u), links[2][0].float().view(C3, 1, 1, C2).permute(0, 3, 1, 2), links[2][1]), acts[2])
    ref = u.permute(0, 2, 3, 1)
    if use_res:
        ref = ref + res.tensor().float()
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 3e-2
    bad = err > tol
    assert not bad.any(), (f"case {case}: {int(bad.sum())} / {bad.numel()} mismatches, max err {float(err.max()):.4f}, "
                           f"first bad index {bad.nonzero()[0].tolist()}, got {got[tuple(bad.nonzero()[0])]:.4f} "
                           f"ref {ref[tuple(bad.nonzero()[0])]:.4f}")


def test_chain_rejects_in_place_and_unsupported():
    from ayolov2_b200 import ops

    x = ops.new_act(1, 16, 16, 64)
    x.buf.zero_()
    links = [ops.pack_chain_weight(torch.zeros(64, 64, 1, 1, device="cuda"), None),
             ops.pack_chain_weight(torch.zeros(64, 64, 3, 3, device="cuda"), None)]
    with pytest.raises(RuntimeError, match="in place"):
        ops.ChainPlan(x, x, links, (1, 1))
    d = ops.chain_desc(x, 256, 64, 0, 1, 1, 0, 64, 0)  # c1 > 128
    assert not ops.chain_supported(d)
    d = ops.chain_desc(x, 64, 64, 0, 1, 1, 0, 64, 0, stride=2)
    assert not ops.chain_supported(d)


def test_offline_factorisation_on_the_device_feeds_the_fused_chain():
    """the offline tool's per-layer work (EVB rank estimate + HOOI, decomposition.py:157-424) with the weight
    resident on the GPU -- torch.linalg on the device, the rank search batched on the host -- gives the same ranks and the
    same chain (as a function: factor signs are not unique) as on the CPU (the decomposed model's eval-mode execution through
    the fused chain kernel is covered by tests/test_model_gpu.py)."""
    import torch.nn as nn

    from ayolov2_b200 import decomposition as dec

    torch.manual_seed(11)
    # a Tucker-structured signal (ranks 12 out / 10 in) + noise, so that EVB truncates both modes
    core, u, v = torch.randn(12, 10, 3, 3), torch.randn(64, 12), torch.randn(64, 10)
    conv = nn.Conv2d(64, 64, 3, padding=1, bias=True)
    with torch.no_grad():
        conv.weight.copy_(torch.einsum("abhw,oa,ib->oihw", core, u, v) / 120 ** 0.5 + 0.05 * torch.randn(64, 64, 3, 3))
    ranks_cpu = dec.estimate_ranks(conv)
    chain_cpu = dec.tucker_decomposition_conv_layer(conv, ranks_cpu)
    conv_gpu = nn.Conv2d(64, 64, 3, padding=1, bias=True).cuda()
    conv_gpu.load_state_dict(conv.state_dict())
    ranks_gpu = dec.estimate_ranks(conv_gpu)
    assert ranks_gpu == ranks_cpu and 0 < ranks_cpu[0] < 64 and 0 < ranks_cpu[1] < 64, (ranks_cpu, ranks_gpu)
    chain_gpu = dec.tucker_decomposition_conv_layer(conv_gpu, ranks_gpu)
    assert all(p.is_cuda for p in chain_gpu.parameters())
    x = torch.randn(2, 64, 24, 24)
    with torch.no_grad():
        want = chain_cpu(x)
        got = chain_gpu(x.cuda()).cpu()
        full = conv(x)
    assert float((got - want).abs().max() / want.abs().max()) < 1e-3
    assert float((want - full).norm() / full.norm()) < 0.2  # the truncation keeps the signal
