"""BottleneckCSP (the block of the reference's tests/res/configs/model_yolov5s_repr.yaml:23-33) on the
CUDA engine vs the fp32 CPU oracle: the bare module (with and without shortcut, several repeat counts) and a whole
detection model whose C3 blocks are all BottleneckCSP (tests/res/yolov5s_csp.yaml). The engine folds the post-concat
BatchNorm + SiLU into the two plain convolutions that feed the concat (engine.py `bottleneck_csp`), so this also checks
that fold against the literal cat -> BN -> act order of the oracle."""
import os

import pytest
import torch

from _parity import errs, record

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _randomize_bn(m, seed):
    g = torch.Generator().manual_seed(seed)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            c = mod.num_features
            mod.weight.data = 1.0 + torch.rand(c, generator=g)
            mod.bias.data = 0.2 * torch.randn(c, generator=g)
            mod.running_mean.data = 0.2 * torch.randn(c, generator=g)
            mod.running_var.data = 0.5 + torch.rand(c, generator=g)


@pytest.mark.parametrize("cin,cout,n,shortcut,hw", [(64, 64, 1, True, (40, 40)), (128, 128, 3, True, (24, 40)),
                                                    (256, 128, 2, False, (20, 20)), (32, 64, 1, True, (80, 48))])
def test_bottleneck_csp_module(cin, cout, n, shortcut, hw):
    from kindle.modules import BottleneckCSP
    from oracle import yolo_oracle

    torch.manual_seed(cin + n)
    m = BottleneckCSP(cin, cout, n_repeat=n, shortcut=shortcut, activation="SiLU").eval()
    _randomize_bn(m, 3)
    x = torch.randn((2, cin, *hw), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = yolo_oracle.bottleneck_csp(m, x)
    got = m.cuda()(x.cuda()).float().cpu()
    assert got.shape == want.shape
    e = errs(got, want)
    record(f"csp_module/{cin}-{cout}-n{n}-{'sc' if shortcut else 'nosc'}", **e)
    assert e["max_norm"] < 1e-2 and e["rel_l2"] < 1e-2, e  # bf16 bound of the north star


def test_bottleneck_csp_model():
    import kindle
    from oracle import yolo_oracle

    torch.manual_seed(0)
    model = kindle.YOLOModel(os.path.join(HERE, "res", "yolov5s_csp.yaml"), verbose=False, init_bias=True).eval()
    assert sum(type(m).__name__ == "BottleneckCSP" for m in model.model) == 8
    _randomize_bn(model, 1)
    x = torch.rand((2, 3, 256, 320), generator=torch.Generator().manual_seed(4))
    want_pred, want_raw = yolo_oracle.forward(model, x)
    got_pred, got_raw = model.cuda()(x.cuda())
    torch.cuda.synchronize()
    for i, (g, w) in enumerate(zip(got_raw, want_raw)):
        e = errs(g, w)
        record(f"csp_model/logits_P{i + 3}", **e)
        assert e["max_norm"] < 1e-2 and e["rel_l2"] < 1e-2, e
