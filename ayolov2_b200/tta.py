"""Test-time augmentation around the CUDA forward: mirror of scripts/utils/tta_utils.py:14-86 and
scripts/utils/torch_utils.py:305-331 (scale_img). Pure orchestration over `model(x)[0]` -- every augmented view goes
through the same sm_100a engine (one engine per distinct input shape, cached by the model)."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def scale_img(img: torch.Tensor, ratio: float = 1.0, same_shape: bool = False, gs: int = 32) -> torch.Tensor:
    """(bs, 3, h, w) scaled by `ratio` (bilinear), padded with 0.447 to a multiple of `gs` (torch_utils.py:305-331)."""
    if ratio == 1.0:
        return img
    h, w = img.shape[2:]
    s = (int(h * ratio), int(w * ratio))
    img = F.interpolate(img, size=s, mode="bilinear", align_corners=False)
    if not same_shape:
        h, w = (math.ceil(x * ratio / gs) * gs for x in (h, w))
    return F.pad(img, [0, w - s[1], 0, h - s[0]], value=0.447)


def descale_pred(p: torch.Tensor, flips: Optional[int], scale: float, img_size: Sequence[int]) -> torch.Tensor:
    """Inverse of the augmentation on the decoded boxes (tta_utils.py:14-36): un-scale, un-flip (2: up-down, 3: left-right)."""
    p[..., :4] /= scale
    if flips == 2:
        p[..., 1] = img_size[0] - p[..., 1]
    elif flips == 3:
        p[..., 0] = img_size[1] - p[..., 0]
    return p


def clip_augmented(model: nn.Module, y: List[torch.Tensor]) -> List[torch.Tensor]:
    """Drop the largest-stride rows of the first view and the smallest-stride rows of the last (tta_utils.py:39-59)."""
    nl = model.model[-1].nl
    g = sum(4 ** x for x in range(nl))
    e = 1
    i = (y[0].shape[1] // g) * sum(4 ** x for x in range(e))
    y[0] = y[0][:, :-i]
    i = (y[-1].shape[1] // g) * sum(4 ** (nl - 1 - x) for x in range(e))
    y[-1] = y[-1][:, i:]
    return y


def inference_with_tta(model: nn.Module, x: torch.Tensor, s: Sequence[float], f: Sequence[Optional[int]]) -> Tuple[torch.Tensor, None]:
    """Reference signature (tta_utils.py:62-86): concatenated, de-augmented predictions of every (scale, flip) view."""
    img_size = x.shape[-2:]
    y = []
    for si, fi in zip(s, f):
        xi = scale_img(x.flip(fi) if fi else x, si, gs=int(model.stride.max()))
        yi = model(xi)[0].clone()  # the engine's prediction buffer is reused by the next view
        y.append(descale_pred(yi, fi, si, img_size))
    y = clip_augmented(model, y)
    return torch.cat(y, 1), None
