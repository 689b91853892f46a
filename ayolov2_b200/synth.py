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


def calibrate_head(model: nn.Module, raw_levels, cand_frac: float = 0.08, conf_thres: float = 0.25, per_level: bool = True) -> None:
    """Random weights give ~0 NMS candidates at conf 0.25, which would make the NMS leg of the
    benchmark vacuous. Shift the head biases (random-init anyway) so that the load looks like the one SURVEY.md §8(d)
    prescribes -- ~cand_frac of the rows (~2,000 of 25,200 per image) are NMS candidates, spread over every pyramid level
    and over many classes:
      * classes: every (anchor, class) bias is set to minus the mean of its observed logit, so no class wins by its random
        offset alone and the arg-max classes spread over the label set;
      * objectness: per level (or globally with per_level=False) the bias shift is solved (bisection on the sample) so that
        exactly cand_frac of the rows pass BOTH tests of metrics.py:313-364: obj > conf_thres and obj * best class > conf_thres.
    `raw_levels`: list of (B, na, ny, nx, no) logits of a sample batch."""
    head = model.model[-1]
    g = torch.Generator().manual_seed(0)

    def sample(r: torch.Tensor):
        r = r.reshape(-1, r.shape[-1]).float().cpu()
        if r.shape[0] > 100000:
            r = r[torch.randperm(r.shape[0], generator=g)[:100000]]
        return r

    def solve(obj: torch.Tensor, best_cls: torch.Tensor) -> float:
        lo, hi = -30.0, 30.0
        for _ in range(50):
            mid = 0.5 * (lo + hi)
            o = torch.sigmoid(obj + mid)
            frac = float(((o > conf_thres) & (o * best_cls > conf_thres)).float().mean())
            lo, hi = (mid, hi) if frac < cand_frac else (lo, mid)
        return 0.5 * (lo + hi)

    samples = []
    with torch.no_grad():
        for conv, r in zip(head.conv, raw_levels):
            b = conv.bias.view(head.na, -1)
            mean_c = r[..., 5:].float().mean(dim=(0, 2, 3))                 # (na, nc)
            b[:, 5:] -= mean_c.to(b.device)
            rs = sample((r.float() - torch.cat((torch.zeros_like(mean_c[:, :5]), mean_c), 1)[None, :, None, None, :].to(r.device)))
            samples.append((rs[:, 4], torch.sigmoid(rs[:, 5:]).max(1).values))
        if per_level:
            for conv, (o, c) in zip(head.conv, samples):
                conv.bias.view(head.na, -1)[:, 4] += solve(o, c)
        else:
            shift = solve(torch.cat([o for o, _ in samples]), torch.cat([c for _, c in samples]))
            for conv in head.conv:
                conv.bias.view(head.na, -1)[:, 4] += shift


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
