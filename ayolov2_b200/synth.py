"""Seeded synthetic model construction for benchmarks, smoke and tests (random-init weights of a named config;
there is no network for checkpoints)."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root


def config_path(name: str = "yolov5s") -> str:
    return os.path.join(ROOT, "res", "configs", "model", f"{name}.yaml")


def build_model(name: str = "yolov5s", seed: int = 0, randomize_bn: bool = True) -> nn.Module:
    """Random-init YOLOModel with non-trivial BatchNorm statistics (so that BN folding is exercised) and the
    detection-prior head biases. Deterministic for a given seed."""
    import kindle

    torch.manual_seed(seed)
    m = kindle.YOLOModel(config_path(name), verbose=False, init_bias=True)
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1)
        for mod in m.modules():
            if isinstance(mod, nn.BatchNorm2d):
                c = mod.num_features
                mod.weight.data = 1.0 + 1.0 * torch.rand(c, generator=g)
                mod.bias.data = 0.2 * torch.randn(c, generator=g)
                mod.running_mean.data = 0.2 * torch.randn(c, generator=g)
                mod.running_var.data = 0.5 + torch.rand(c, generator=g)
    return m.eval()


def calibrate_head(model: nn.Module, raw_levels, cand_frac: float = 0.08, obj_level: float = 0.3) -> None:
    """Random weights give ~0 NMS candidates at conf 0.25, which would make the NMS leg of the
    benchmark vacuous. Shift the head biases (random-init anyway) so that ~cand_frac of the 25,200 rows per image
    carry objectness > obj_level and class scores are O(1): objectness bias += logit(obj_level) - q_{1-cand_frac}
    of the observed objectness logits, class biases reset to 0. `raw_levels`: list of (B, na, ny, nx, no) logits."""
    import math

    head = model.model[-1]
    obj = torch.cat([r[..., 4].reshape(-1).float().cpu() for r in raw_levels])
    q = torch.quantile(obj[torch.randperm(obj.numel())[:200000]], 1.0 - cand_frac).item()
    shift = math.log(obj_level / (1.0 - obj_level)) - q
    with torch.no_grad():
        for conv in head.conv:
            b = conv.bias.view(head.na, -1)
            b[:, 4] += shift
            b[:, 5:] = 0.0
