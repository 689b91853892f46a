"""kindle.modules.bottleneck — Bottleneck, C3, BottleneckCSP.

Child names follow the pickled fixture: C3 has conv1, conv2, conv3, bottleneck_c3 (Sequential of Bottleneck
with conv1 1x1 / conv2 3x3 and a `shortcut` flag); hidden width = 0.5 * out_channels.
"""
from __future__ import annotations

from typing import Union

import torch
import torch.nn as nn

from .activation import Activation
from .conv import Conv


class Bottleneck(nn.Module):
    """x + conv2_3x3(conv1_1x1(x)) when `shortcut` and in == out channels."""

    def __init__(self, in_channels: int, out_channels: int, shortcut: bool = True, groups: int = 1,
                 expansion: float = 0.5, activation: Union[str, None] = "ReLU") -> None:
        super().__init__()
        hidden = int(out_channels * expansion)
        self.conv1 = Conv(in_channels, hidden, 1, 1, activation=activation)
        self.conv2 = Conv(hidden, out_channels, 3, 1, groups=groups, activation=activation)
        self.shortcut = bool(shortcut) and in_channels == out_channels

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ...engine import run_single_module

        return run_single_module(self, x)


class C3(nn.Module):
    """conv3(cat[bottlenecks(conv1(x)), conv2(x)], 1). yaml: `[-1, n, C3, [out_channels(, shortcut)], {...}]`."""

    def __init__(self, in_channels: int, out_channels: int, n_repeat: int = 1, shortcut: bool = True, groups: int = 1,
                 expansion: float = 0.5, activation: Union[str, None] = "ReLU") -> None:
        super().__init__()
        hidden = int(out_channels * expansion)
        self.conv1 = Conv(in_channels, hidden, 1, 1, activation=activation)
        self.conv2 = Conv(in_channels, hidden, 1, 1, activation=activation)
        self.conv3 = Conv(2 * hidden, out_channels, 1, 1, activation=activation)
        self.bottleneck_c3 = nn.Sequential(*[
            Bottleneck(hidden, hidden, shortcut=shortcut, groups=groups, expansion=1.0, activation=activation)
            for _ in range(n_repeat)
        ])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ...engine import run_single_module

        return run_single_module(self, x)


class BottleneckCSP(nn.Module):
    """CSP bottleneck (ultralytics v3/v4 form): conv4(act(bn(cat[conv3(m(conv1 x)), conv2(x)]))), with plain
    (bias-free, no BN) conv2/conv3. Appears only in tests/res/configs/model_yolov5s_repr.yaml:23-33; there is no
    reference fixture for it, so child names here are this package's own ("not verifiable")."""

    def __init__(self, in_channels: int, out_channels: int, n_repeat: int = 1, shortcut: bool = True, groups: int = 1,
                 expansion: float = 0.5, activation: Union[str, None] = "ReLU") -> None:
        super().__init__()
        hidden = int(out_channels * expansion)
        self.conv1 = Conv(in_channels, hidden, 1, 1, activation=activation)
        self.conv2 = nn.Conv2d(in_channels, hidden, 1, 1, bias=False)
        self.conv3 = nn.Conv2d(hidden, hidden, 1, 1, bias=False)
        self.conv4 = Conv(2 * hidden, out_channels, 1, 1, activation=activation)
        self.batch_norm = nn.BatchNorm2d(2 * hidden, eps=1e-3, momentum=0.03)
        self.activation = Activation(activation)()
        self.bottleneck_csp = nn.Sequential(*[
            Bottleneck(hidden, hidden, shortcut=shortcut, groups=groups, expansion=1.0, activation=activation)
            for _ in range(n_repeat)
        ])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ...engine import run_single_module

        return run_single_module(self, x)
