"""CPU oracle for the model forward: a plain PyTorch fp32 restatement of the kindle operators.
TEST INFRASTRUCTURE ONLY — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (ayolov2_b200/) never does.

`kindle` (PyPI, JeiKeiLim, `>=0.4.12` per environment.yml:42, not vendored under /root/reference and not
installable here) is the third-party dependency that holds these operators; its algorithm is restated from
  * the yaml graphs                         res/configs/model/*.yaml
  * the pickled module tree of the fixture  tests/res/weights/yolov5s_kindle.pt (child names, Conv2d/BN hyper-params,
                                            head buffers)
  * the reference's call sites              scripts/utils/train_utils.py:436-444 (eval output tuple),
                                            scripts/loss/losses.py:245-255 (train-layout tensors, xy/wh decode)
  * the published YOLOv5 operator definitions those yaml rows name.
Parity pins (tests/test_oracle_model.py): parameter counts of README.md:206-211 and tests/test_tensor_decomposition.py:47,
the 45/132 freeze split of tests/test_model_manager.py:60-61, fuse invariance (tests/test_model_convert.py:43-44) and
the fixture-weights -> detections pin on the reference's own 99 COCO val images (608 detections /
411 TP@0.5). Forward *numerics* have no reference golden vectors: "parity unpinned" beyond those behavioural pins.

The functions walk any module tree that uses kindle's child names, so they run on this repo's kindle-compatible
classes and on the reference's pickles alike.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F


# When set, the forward rounds to bfloat16 (straight-through: gradients pass unchanged) at exactly the points where the
# CUDA path stores bf16: conv weights, conv inputs, the raw conv output, and each layer output. Train-mode BatchNorm
# re-normalises with batch statistics at every layer, which makes a random-init network chaotic w.r.t. such rounding
# (the rounded and the exact fp32 forward give parameter gradients with cosine ~0.78 on the small test inputs), so
# training-step parity is measured against this rounding-matched oracle, and the rounding noise itself is reported.
SIMULATE_BF16 = False


def _r(t: torch.Tensor) -> torch.Tensor:
    if not SIMULATE_BF16:
        return t
    return t + (t.to(torch.bfloat16).float() - t).detach()


def _act(m: nn.Module, x: torch.Tensor) -> torch.Tensor:
    a = getattr(m, "activation", None)
    if a is None or isinstance(a, nn.Identity):
        return x
    return a(x)


def conv_bn_act(m: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """kindle Conv: activation(batch_norm(conv(x))); `conv` may be a Tucker nn.Sequential
    (scripts/tensor_decomposition/decomposition.py:325-335); `batch_norm` is Identity after fuse()."""
    if SIMULATE_BF16 and isinstance(m.conv, nn.Conv2d):
        c = m.conv
        y = _r(F.conv2d(_r(x), _r(c.weight), c.bias, c.stride, c.padding, c.dilation, c.groups))
    else:
        y = m.conv(x)
    bn = getattr(m, "batch_norm", None)
    if isinstance(bn, nn.BatchNorm2d):
        y = bn(y)
    return _act(m, y)


def focus(m: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """Slice order verified by the detection pin."""
    x = torch.cat((x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]), 1)
    return _r(conv_bn_act(m, x))


def bottleneck(m: nn.Module, x: torch.Tensor) -> torch.Tensor:
    t = _r(conv_bn_act(m.conv1, x))
    y = conv_bn_act(m.conv2, t)
    return _r(x + y) if m.shortcut else _r(y)


def c3(m: nn.Module, x: torch.Tensor) -> torch.Tensor:
    y1 = _r(conv_bn_act(m.conv1, x))
    for b in m.bottleneck_c3:
        y1 = bottleneck(b, y1)
    return _r(conv_bn_act(m.conv3, torch.cat((y1, _r(conv_bn_act(m.conv2, x))), 1)))


def bottleneck_csp(m: nn.Module, x: torch.Tensor) -> torch.Tensor:
    y1 = conv_bn_act(m.conv1, x)
    for b in m.bottleneck_csp:
        y1 = bottleneck(b, y1)
    y1 = m.conv3(y1)
    y2 = m.conv2(x)
    return conv_bn_act(m.conv4, m.activation(m.batch_norm(torch.cat((y1, y2), 1))))


def spp(m: nn.Module, x: torch.Tensor) -> torch.Tensor:
    x1 = _r(conv_bn_act(m.conv1, x))
    return _r(conv_bn_act(m.conv2, torch.cat([x1] + [p(x1) for p in m.pooling_modules], 1)))


def sppf(m: nn.Module, x: torch.Tensor) -> torch.Tensor:
    x1 = _r(conv_bn_act(m.conv1, x))
    p1 = m.pooling(x1)
    p2 = m.pooling(p1)
    p3 = m.pooling(p2)
    return _r(conv_bn_act(m.conv2, torch.cat((x1, p1, p2, p3), 1)))


def yolo_head(m: nn.Module, xs: Sequence[torch.Tensor], training: bool):
    """Train: [(bs, na, ny, nx, no)] logits. Eval: (cat over levels of (bs, na*ny*nx, no), [logits...])."""
    raw, dec = [], []
    for i, (conv, x) in enumerate(zip(m.conv, xs)):
        t = _r(F.conv2d(_r(x), _r(conv.weight), conv.bias)) if SIMULATE_BF16 else conv(x)
        bs, _, ny, nx = t.shape
        t = t.view(bs, m.na, m.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        raw.append(t)
        if not training:
            yv, xv = torch.meshgrid(torch.arange(ny), torch.arange(nx), indexing="ij")
            grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).to(t)  # dtype and device (GPU-run checks at full size)
            y = t.sigmoid()
            stride = float(m.stride[i])
            xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * stride
            wh = (y[..., 2:4] * 2) ** 2 * m.anchor_grid[i].to(t.dtype).view(1, m.na, 1, 1, 2)
            if getattr(m, "out_xyxy", False):
                xy, wh = xy - wh / 2, xy + wh / 2
            dec.append(torch.cat((xy, wh, y[..., 4:]), -1).view(bs, -1, m.no))
    return raw if training else (torch.cat(dec, 1), raw)


def _conv_layer(m: nn.Module, x: torch.Tensor) -> torch.Tensor:
    return _r(conv_bn_act(m, x))


_DISPATCH = {
    "Conv": _conv_layer, "Focus": focus, "Bottleneck": bottleneck, "C3": c3, "BottleneckCSP": bottleneck_csp,
    "SPP": spp, "SPPF": sppf,
}


def layer_forward(m: nn.Module, x, training: bool):
    name = type(m).__name__
    if name in _DISPATCH:
        return _DISPATCH[name](m, x)
    if name == "Concat":
        return torch.cat(list(x), getattr(m, "dimension", 1))
    if name == "Upsample":
        return F.interpolate(x, scale_factor=m.scale_factor, mode=m.mode)
    if name == "YOLOHead":
        return yolo_head(m, x, training)
    raise NotImplementedError(name)


@torch.no_grad()
def forward(model: nn.Module, x: torch.Tensor, training: bool = False, return_layers: bool = False):
    """YOLOModel.forward restated: walk `model.model`, feeding each layer from its `from_idx`."""
    outs: List[Union[torch.Tensor, tuple, list]] = []
    x = x.float()
    for i, m in enumerate(model.model):
        frm = getattr(m, "from_idx", -1)
        if isinstance(frm, (list, tuple)):
            inp = [x if f == -1 else outs[f] for f in frm]
        else:
            inp = x if frm == -1 else outs[frm]
        x = layer_forward(m, inp, training)
        outs.append(x)
    return (x, outs) if return_layers else x


def forward_with_grad(model: nn.Module, x: torch.Tensor):
    """Training-mode forward with autograd enabled (BN in its current mode); used by the loss/backward parity tests."""
    outs: List = []
    for i, m in enumerate(model.model):
        frm = getattr(m, "from_idx", -1)
        if isinstance(frm, (list, tuple)):
            inp = [x if f == -1 else outs[f] for f in frm]
        else:
            inp = x if frm == -1 else outs[frm]
        x = layer_forward(m, inp, True)
        outs.append(x)
    return x
