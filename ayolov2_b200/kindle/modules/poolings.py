"""kindle.modules.poolings — SPP and SPPF. Hidden width = in_channels // 2."""
from __future__ import annotations

from typing import Sequence, Union

import torch
import torch.nn as nn

from .conv import Conv


class SPP(nn.Module):
    """conv2(cat[x1, p5(x1), p9(x1), p13(x1)]), x1 = conv1(x). yaml: `[-1, 1, SPP, [c, [5, 9, 13]], {...}]`."""

    def __init__(self, in_channels: int, out_channels: int, kernel_sizes: Sequence[int] = (5, 9, 13),
                 activation: Union[str, None] = "ReLU") -> None:
        super().__init__()
        hidden = in_channels // 2
        self.conv1 = Conv(in_channels, hidden, 1, 1, activation=activation)
        self.conv2 = Conv(hidden * (len(kernel_sizes) + 1), out_channels, 1, 1, activation=activation)
        self.pooling_modules = nn.ModuleList([nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)
                                              for k in kernel_sizes])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ...engine import run_single_module

        return run_single_module(self, x)


class SPPF(nn.Module):
    """conv2(cat[x1, p(x1), p(p(x1)), p(p(p(x1)))]), p = MaxPool2d(k, 1, k//2). yaml: `[-1, 1, SPPF, [c, 5], {...}]`
    (res/configs/model/yolov5s.yaml:33)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 5,
                 activation: Union[str, None] = "ReLU") -> None:
        super().__init__()
        hidden = in_channels // 2
        self.conv1 = Conv(in_channels, hidden, 1, 1, activation=activation)
        self.conv2 = Conv(hidden * 4, out_channels, 1, 1, activation=activation)
        self.pooling = nn.MaxPool2d(kernel_size=kernel_size, stride=1, padding=kernel_size // 2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ...engine import run_single_module

        return run_single_module(self, x)
