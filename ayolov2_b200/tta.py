"""Test-time augmentation around the CUDA forward.

Behavioural mirror of the reference's augmented inference (scripts/utils/tta_utils.py:14-86 with the image rescaling of
scripts/utils/torch_utils.py:305-331): each view is the input flipped and / or rescaled to a stride multiple, goes through
`model(x)[0]` -- i.e. the same sm_100a engine, one engine per distinct input shape -- and its decoded boxes are mapped back
to the un-augmented image before all views are concatenated. This module contains no kernels.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

PAD_VALUE = 0.447  # the reference pads rescaled views with the ImageNet mean
FLIP_UD, FLIP_LR = 2, 3  # tensor dimensions of an NCHW batch


@dataclass(frozen=True)
class View:
    """One augmentation: isotropic scale factor and an optional flip axis (2 = up-down, 3 = left-right)."""

    scale: float
    flip: Optional[int]

    def apply(self, x: torch.Tensor, stride: int) -> torch.Tensor:
        return scale_img(x.flip(self.flip) if self.flip else x, self.scale, gs=stride)

    def undo(self, pred: torch.Tensor, hw: Sequence[int]) -> torch.Tensor:
        """Boxes (cx, cy, w, h in view pixels) back to the original image, in place."""
        pred[..., :4] /= self.scale
        if self.flip == FLIP_UD:
            pred[..., 1] = hw[0] - pred[..., 1]
        elif self.flip == FLIP_LR:
            pred[..., 0] = hw[1] - pred[..., 0]
        return pred


def scale_img(img: torch.Tensor, ratio: float = 1.0, same_shape: bool = False, gs: int = 32) -> torch.Tensor:
    """Bilinear rescale of an NCHW batch by `ratio`; unless `same_shape`, the canvas grows to the next multiple of `gs`."""
    if ratio == 1.0:
        return img
    src_h, src_w = img.shape[2:]
    new_h, new_w = int(src_h * ratio), int(src_w * ratio)
    out = F.interpolate(img, size=(new_h, new_w), mode="bilinear", align_corners=False)
    canvas_h, canvas_w = (src_h, src_w) if same_shape else tuple(math.ceil(v * ratio / gs) * gs for v in (src_h, src_w))
    return F.pad(out, [0, canvas_w - new_w, 0, canvas_h - new_h], value=PAD_VALUE)


def descale_pred(p: torch.Tensor, flips: Optional[int], scale: float, img_size: Sequence[int]) -> torch.Tensor:
    """Reference-named entry point for View.undo."""
    return View(scale, flips).undo(p, img_size)


def clip_augmented(model: nn.Module, y: List[torch.Tensor]) -> List[torch.Tensor]:
    """The first view loses the rows of its coarsest detection level, the last view those of its finest: with nl levels
    whose row counts relate as 4^(nl-1) : ... : 4 : 1, the coarsest level is 1 / sum(4^k) of a view's rows."""
    nl = model.model[-1].nl
    unit = sum(4 ** k for k in range(nl))
    coarse_rows = y[0].shape[1] // unit             # rows of the stride-max level in the first view
    fine_rows = (y[-1].shape[1] // unit) * 4 ** (nl - 1)  # rows of the stride-min level in the last view
    y[0] = y[0][:, :-coarse_rows]
    y[-1] = y[-1][:, fine_rows:]
    return y


def inference_with_tta(model: nn.Module, x: torch.Tensor, s: Sequence[float], f: Sequence[Optional[int]]) -> Tuple[torch.Tensor, None]:
    """`(cat of the de-augmented predictions of every (scale, flip) view, None)` -- the reference's return convention."""
    hw = x.shape[-2:]
    stride = int(model.stride.max())
    outs = []
    for view in (View(si, fi) for si, fi in zip(s, f)):
        pred = model(view.apply(x, stride))[0].clone()  # the engine's prediction buffer is reused by the next view
        outs.append(view.undo(pred, hw))
    return torch.cat(clip_augmented(model, outs), 1), None
