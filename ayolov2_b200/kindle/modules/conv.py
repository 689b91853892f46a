"""kindle.modules.conv — Conv (conv -> BatchNorm(eps 1e-3, momentum 0.03) -> activation) and Focus.

API recovered from the reference's call sites and pickled fixture:
children `conv` (nn.Conv2d, bias=False, padding k//2 unless given), `batch_norm`, `activation`;
after `YOLOModel.fuse()` the BN is folded into `conv` (bias added) and `batch_norm` becomes Identity
(val.py:331, tests/test_model_convert.py:43-44).

These are parameter containers plus graph descriptions: execution happens in ayolov2_b200.engine
(CUDA kernels); `forward` on a bare module runs a one-module engine on CUDA tensors.
"""
from __future__ import annotations

from typing import Optional, Union

import torch
import torch.nn as nn

from .activation import Activation


def autopad(kernel_size: int, padding: Optional[int] = None) -> int:
    return kernel_size // 2 if padding is None else padding


class Conv(nn.Module):
    """yaml: `[-1, 1, Conv, [out_channels, kernel, stride(, padding)], {activation: SiLU}]`."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 1, stride: int = 1,
                 padding: Optional[int] = None, groups: int = 1, activation: Union[str, None] = "ReLU") -> None:
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = kernel_size, stride
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, autopad(kernel_size, padding),
                              groups=groups, bias=False)
        self.batch_norm = nn.BatchNorm2d(out_channels, eps=1e-3, momentum=0.03)
        self.activation = Activation(activation)()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ...engine import run_single_module

        return run_single_module(self, x)

    def fusefoward(self, x: torch.Tensor) -> torch.Tensor:  # kindle spelling, kept for API compatibility
        return self.forward(x)


class Focus(nn.Module):
    """Space-to-depth then Conv: cat[x[::2, ::2], x[1::2, ::2], x[::2, 1::2], x[1::2, 1::2]] on (H, W).

    yaml: `[-1, 1, Focus, [out_channels, kernel], {activation: SiLU}]` (res/configs/model/yolov5_v5.yaml:21).
    The pickled fixture keeps conv/batch_norm/activation directly on the Focus module."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 1, stride: int = 1,
                 padding: Optional[int] = None, groups: int = 1, activation: Union[str, None] = "ReLU") -> None:
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = kernel_size, stride
        self.conv = nn.Conv2d(in_channels * 4, out_channels, kernel_size, stride, autopad(kernel_size, padding),
                              groups=groups, bias=False)
        self.batch_norm = nn.BatchNorm2d(out_channels, eps=1e-3, momentum=0.03)
        self.activation = Activation(activation)()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ...engine import run_single_module

        return run_single_module(self, x)
