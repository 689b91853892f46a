"""kindle.modules.yolo_head — YOLOHead.

Attributes the reference touches: nc, no, nl, na (+ n_classes, n_outputs, n_layers, n_anchors), out_xyxy, stride,
buffers `anchors` (nl, na, 2) in grid units and `anchor_grid` (nl, 1, na, 1, 1, 2) in pixels, `conv` ModuleList
of 1x1 Conv2d with bias (scripts/loss/losses.py:201-221, scripts/utils/anchors.py:28-36,204-228, export.py:171).
Train mode returns [(bs, na, ny, nx, no)] raw logits; eval mode returns (cat(bs, sum na*ny*nx, no), [raw...])
with y = sigmoid, xy = (y*2 - 0.5 + grid) * stride, wh = (y*2)^2 * anchor_grid.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch
import torch.nn as nn


class YOLOHead(nn.Module):
    def __init__(self, in_channels: Sequence[int], n_classes: int, anchors: Sequence[Sequence[float]],
                 out_xyxy: bool = False) -> None:
        super().__init__()
        self.out_xyxy = out_xyxy
        self.n_classes = self.nc = n_classes
        self.n_outputs = self.no = n_classes + 5
        self.n_layers = self.nl = len(anchors)
        self.n_anchors = self.na = len(anchors[0]) // 2
        self.grid = [torch.zeros(1)] * self.nl
        self.stride = torch.zeros(self.nl)
        a = torch.tensor(anchors).float().view(self.nl, -1, 2)
        self.register_buffer("anchors", a)  # divided by stride once the strides are known (YOLOModel.__init__)
        self.register_buffer("anchor_grid", a.clone().view(self.nl, 1, -1, 1, 1, 2))
        self.conv = nn.ModuleList(nn.Conv2d(c, self.no * self.na, 1) for c in in_channels)

    def initialize_biases(self, class_frequency=None) -> None:
        """ultralytics-style prior: objectness assumes ~8 objects per 640 image, classes ~0.6/(nc-0.99)."""
        for conv, s in zip(self.conv, self.stride):
            b = conv.bias.view(self.na, -1)
            b.data[:, 4] += math.log(8 / (640 / float(s)) ** 2)
            b.data[:, 5:] += math.log(0.6 / (self.nc - 0.99)) if class_frequency is None else torch.log(
                class_frequency / class_frequency.sum())
            conv.bias = nn.Parameter(b.view(-1), requires_grad=True)

    def forward(self, x: List[torch.Tensor]):
        from ...engine import run_single_module

        return run_single_module(self, x)
