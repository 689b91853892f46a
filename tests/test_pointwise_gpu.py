"""GPU parity of the data-movement kernels against plain PyTorch fp32 references."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_space_to_depth(dtype):
    from ayolov2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    B, H, W = 3, 64, 96
    if dtype == torch.uint8:
        img = torch.randint(0, 256, (B, 3, H, W), device="cuda", generator=g, dtype=torch.uint8)
        scale = 1.0 / 255.0
    else:
        img = torch.rand((B, 3, H, W), device="cuda", generator=g)
        scale = 1.0
    out = ops.new_act(B, H // 2, W // 2, 16)
    ops.space_to_depth(img, out, scale)
    torch.cuda.synchronize()
    x = img.float() * scale
    ref = torch.zeros((B, H // 2, W // 2, 16), device="cuda")
    for dy in range(2):
        for dx in range(2):
            for c in range(3):
                ref[..., (dy * 2 + dx) * 3 + c] = x[:, c, dy::2, dx::2]
    assert torch.equal(out.buf.float(), ref.to(torch.bfloat16).float())


@pytest.mark.parametrize("hw", [(20, 20), (13, 17), (40, 40)])
def test_sppf_pool(hw):
    from ayolov2_b200 import ops

    H, W = hw
    g = torch.Generator(device="cuda").manual_seed(0)
    B, C_ = 3, 32
    buf = torch.randn((B, H, W, 4 * C_), device="cuda", generator=g).to(torch.bfloat16)
    v = ops.ActView(buf, 0, 4 * C_)
    ops.sppf_pool(v.slice(0, C_), v.slice(C_, C_), v.slice(2 * C_, C_), v.slice(3 * C_, C_), (5, 9, 13))
    torch.cuda.synchronize()
    x = buf[..., :C_].float().permute(0, 3, 1, 2)
    for i, k in enumerate((5, 9, 13)):
        ref = F.max_pool2d(x, k, 1, k // 2).permute(0, 2, 3, 1)
        assert torch.equal(buf[..., (i + 1) * C_:(i + 2) * C_].float(), ref), k
    # SPPF cascade identity: p(p(x)) == 9-window, p(p(p(x))) == 13-window
    p1 = F.max_pool2d(x, 5, 1, 2)
    assert torch.equal(F.max_pool2d(p1, 5, 1, 2), F.max_pool2d(x, 9, 1, 4))


def test_upsample2x():
    from ayolov2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    B, H, W, C_ = 2, 10, 12, 64
    x = ops.ActView(torch.randn((B, H, W, C_ + 16), device="cuda", generator=g).to(torch.bfloat16), 8, C_)
    ybuf = torch.zeros((B, 2 * H, 2 * W, 2 * C_), device="cuda", dtype=torch.bfloat16)
    y = ops.ActView(ybuf, C_, C_)
    ops.upsample2x(x, y)
    torch.cuda.synchronize()
    ref = F.interpolate(x.tensor().float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y.tensor().float(), ref)
    assert torch.all(ybuf[..., :C_] == 0)


def test_head_decode():
    from ayolov2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    B, ny, nx, na, no = 2, 20, 12, 3, 85
    logits = ops.ActView((torch.randn((B, ny, nx, 256), device="cuda", generator=g) * 2).to(torch.bfloat16), 0, 256)
    anchors = torch.tensor([[116., 90.], [156., 198.], [373., 326.]], device="cuda")
    total = 1000 + na * ny * nx
    pred = torch.zeros((B, total, no), device="cuda")
    raw = torch.zeros((B, na, ny, nx, no), device="cuda")
    ops.head_decode(logits, na, no, 32.0, anchors.reshape(-1).contiguous(), pred, 1000, raw)
    torch.cuda.synchronize()
    t = logits.buf[..., :na * no].float().view(B, ny, nx, na, no).permute(0, 3, 1, 2, 4)
    assert torch.equal(raw, t.contiguous())
    y = torch.sigmoid(t)
    yv, xv = torch.meshgrid(torch.arange(ny, device="cuda"), torch.arange(nx, device="cuda"), indexing="ij")
    grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()
    xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * 32.0
    wh = (y[..., 2:4] * 2) ** 2 * anchors.view(1, na, 1, 1, 2)
    ref = torch.cat((xy, wh, y[..., 4:]), -1).reshape(B, -1, no)
    got = pred[:, 1000:]
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), float((got - ref).abs().max())
    assert torch.all(pred[:, :1000] == 0)
