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


def calibrate_head_bias_only(model: nn.Module, raw_levels, cand_frac: float = 0.08, obj_level: float = 0.3) -> None:
    """The round-1 calibration, kept for the parity tests: only BIASES move (objectness bias += logit(obj_level) - the
    (1 - cand_frac) quantile of the observed objectness logits over all levels, class biases = 0), so the head stays as
    well conditioned as the random initialisation made it and bf16-vs-fp32 comparisons of its logits remain meaningful.
    Its load is lopsided (the candidates are the stride-32 rows, in a handful of classes); the benchmark uses
    `calibrate_head`. `raw_levels`: list of (B, na, ny, nx, no) logits of a sample batch."""
    head = model.model[-1]
    obj = torch.cat([r[..., 4].reshape(-1).float().cpu() for r in raw_levels])
    q = torch.quantile(obj[torch.randperm(obj.numel(), generator=torch.Generator().manual_seed(0))[:200000]], 1.0 - cand_frac).item()
    shift = math.log(obj_level / (1.0 - obj_level)) - q
    with torch.no_grad():
        for conv in head.conv:
            b = conv.bias.view(head.na, -1)
            b[:, 4] += shift
            b[:, 5:] = 0.0
    if hasattr(model, "invalidate_engine"):
        model.invalidate_engine()


def calibrate_head(model: nn.Module, run_raw, cand_frac: float = 0.08, conf_thres: float = 0.25, per_level: bool = True,
                   cls_gain: float = 4.0, refine: int = 2) -> None:
    """Random weights give ~0 NMS candidates at conf 0.25, which would make the NMS leg of the
    benchmark vacuous, and their head logits barely vary over the image (objectness: -6.6 +- 0.003 per anchor at stride 8,
    a tenth of a bf16 step at that magnitude). The detect convolutions (random-init anyway) are therefore re-scaled so that
    the load looks like the one SURVEY.md §8(d) prescribes -- ~cand_frac of the rows (~2,000 of 25,200 per image) are NMS
    candidates, on every pyramid level, over many classes:
      * the objectness / class biases are zeroed and the sample batch is run again, so that the logits are the pure filter
        responses (small numbers, well resolved in bf16);
      * objectness and class filters of every anchor are standardised on that sample: logit' = (logit - mean) / std, i.e.
        weight /= std, bias = -mean / std, so the logits have unit spread (and every class the same); the class logits
        are then stretched by `cls_gain`, so that -- like in a trained detector -- a row's best class score is close to 1
        and the rows that pass the objectness test are (nearly) the candidates (with unit-spread class logits most rows
        that pass objectness fail on the class score: 57 % of the stride-8 rows passed objectness for 8.6 % candidates);
      * the objectness bias then gets the shift (bisection on the sample, per level or globally) at which exactly cand_frac
        of the rows pass BOTH tests of metrics.py:313-364: obj > conf_thres and obj * best class score > conf_thres.
      * `refine` times the sample is run through the model AS IT NOW IS (bf16 weights, re-scaled filters) and the
        objectness biases are re-solved on those logits: the re-scaled filters amplify rounding noise, so the first
        solution, found on the re-scaled OLD logits, overshoots (measured: 17.6 % of the rows instead of the requested
        12 %). Only biases move in these passes, so they converge at once.
    The standardised filters amplify the (tiny) position-dependent part of the random features ~100x, rounding noise
    included: fine for a synthetic LOAD, useless for bf16-vs-fp32 logit comparisons (use calibrate_head_bias_only there).
    `run_raw`: callable returning the list of (B, na, ny, nx, no) logits of a sample batch with the model's CURRENT weights."""
    head = model.model[-1]
    g = torch.Generator().manual_seed(0)

    def refresh():
        if hasattr(model, "invalidate_engine"):
            model.invalidate_engine()

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
        for conv in head.conv:
            conv.bias.view(head.na, -1)[:, 4:] = 0.0
        refresh()
        raw_levels = [r.detach().float() for r in run_raw()]
        for conv, r in zip(head.conv, raw_levels):
            mean = r[..., 4:].mean(dim=(0, 2, 3))                        # (na, 1 + nc)
            std = r[..., 4:].std(dim=(0, 2, 3)).clamp_min(1e-12)
            w = conv.weight.view(head.na, -1, *conv.weight.shape[1:])    # (na, no, cin, 1, 1)
            gain = torch.ones_like(std)
            gain[:, 1:] = cls_gain
            w[:, 4:] *= (gain / std).to(w.device)[:, :, None, None, None]
            conv.bias.view(head.na, -1)[:, 4:] = (-mean / std * gain).to(conv.bias.device)
            z = ((r[..., 4:] - mean[None, :, None, None, :]) / std[None, :, None, None, :] * gain[None, :, None, None, :]
                 ).reshape(-1, r.shape[-1] - 4).cpu()
            if z.shape[0] > 100000:
                z = z[torch.randperm(z.shape[0], generator=g)[:100000]]
            samples.append((z[:, 0], torch.sigmoid(z[:, 1:]).max(1).values))
        if per_level:
            for conv, (o, c) in zip(head.conv, samples):
                conv.bias.view(head.na, -1)[:, 4] += solve(o, c)
        else:
            shift = solve(torch.cat([o for o, _ in samples]), torch.cat([c for _, c in samples]))
            for conv in head.conv:
                conv.bias.view(head.na, -1)[:, 4] += shift
        for _ in range(refine):
            refresh()
            now = []
            for r in run_raw():
                z = r.detach().float()[..., 4:].reshape(-1, r.shape[-1] - 4).cpu()
                now.append((z[:, 0], torch.sigmoid(z[:, 1:]).max(1).values))
            if per_level:
                for conv, (o, c) in zip(head.conv, now):
                    conv.bias.view(head.na, -1)[:, 4] += solve(o, c)
            else:
                shift = solve(torch.cat([o for o, _ in now]), torch.cat([c for _, c in now]))
                for conv in head.conv:
                    conv.bias.view(head.na, -1)[:, 4] += shift
    refresh()


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
