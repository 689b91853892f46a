"""Seeded synthetic model construction for benchmarks, smoke and tests (random-init weights of a named config;
there is no network for checkpoints)."""
from __future__ import annotations

import math
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


def calibrate_head(model: nn.Module, raw_levels, cand_frac: float = 0.08, obj_level: float = 0.3, per_level: bool = True) -> None:
    """Random weights give ~0 NMS candidates at conf 0.25, which would make the NMS leg of the
    benchmark vacuous. Shift the head biases (random-init anyway) so that the load looks like a trained detector's
    (~2,000 candidates per image over many classes and all pyramid levels):
      * objectness: per level (per_level=True) or globally, bias += logit(obj_level) - q_{1-cand_frac} of the observed
        objectness logits, so that ~cand_frac of the rows of EVERY level carry objectness > obj_level;
      * classes: every (anchor, class) bias is set to minus the mean of its observed logit, so no class wins by its
        random offset alone and the arg-max classes spread over the whole label set.
    `raw_levels`: list of (B, na, ny, nx, no) logits of a sample batch."""
    import math

    head = model.model[-1]
    g = torch.Generator().manual_seed(0)

    def quantile(v: torch.Tensor) -> float:
        v = v.reshape(-1).float().cpu()
        if v.numel() > 200000:
            v = v[torch.randperm(v.numel(), generator=g)[:200000]]
        return torch.quantile(v, 1.0 - cand_frac).item()

    target = math.log(obj_level / (1.0 - obj_level))
    q_all = quantile(torch.cat([r[..., 4].reshape(-1).float().cpu() for r in raw_levels]))
    with torch.no_grad():
        for conv, r in zip(head.conv, raw_levels):
            b = conv.bias.view(head.na, -1)
            b[:, 4] += target - (quantile(r[..., 4]) if per_level else q_all)
            if per_level:
                b[:, 5:] -= r[..., 5:].float().mean(dim=(0, 2, 3)).to(b.device)
            else:
                b[:, 5:] = 0.0


def synth_predictions(batch: int, n: int = 25200, nc: int = 80, seed: int = 0, cand_frac: float = 0.08, clusters: int = 200,
                      img: float = 640.0, device="cuda") -> torch.Tensor:
    """The synthetic (batch, n, 5 + nc) prediction tensor SURVEY.md §8(d) prescribes for NMS timing: objectness U(0, 0.2)
    except a Bernoulli(cand_frac) subset at U(0.25, 1) (~2,000 candidates / image at n = 25,200), class probabilities
    U(0, 1) with one class per row boosted to U(0.5, 1), boxes = `clusters` centres per image U(0, img)^2 + N(0, 4 px),
    wh LogU(16, 256) per cluster (jittered) -- real suppression chains. Generated on `device`."""
    g = torch.Generator(device=device).manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g, device=device)  # noqa: E731
    pred = torch.empty((batch, n, 5 + nc), device=device)
    cls = r(batch, n, nc)
    boost = torch.randint(0, nc, (batch, n, 1), generator=g, device=device)
    cls.scatter_(2, boost, 0.5 + 0.5 * r(batch, n, 1))
    obj = torch.where(r(batch, n) < cand_frac, 0.25 + 0.75 * r(batch, n), 0.2 * r(batch, n))
    centres = r(batch, clusters, 2) * img
    which = torch.randint(0, clusters, (batch, n), generator=g, device=device)
    idx = which[..., None].expand(-1, -1, 2)
    xy = torch.gather(centres, 1, idx) + 4.0 * torch.randn((batch, n, 2), generator=g, device=device)
    cwh = torch.exp(math.log(16.0) + (math.log(256.0) - math.log(16.0)) * r(batch, clusters, 2))
    wh = torch.gather(cwh, 1, idx) * torch.exp(0.1 * torch.randn((batch, n, 2), generator=g, device=device))
    pred[..., :2], pred[..., 2:4], pred[..., 4], pred[..., 5:] = xy, wh, obj, cls
    return pred
