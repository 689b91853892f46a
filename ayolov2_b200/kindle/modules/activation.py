"""Activation lookup used by the kindle-compatible modules (yaml kwarg `activation: SiLU`)."""
from __future__ import annotations

from typing import Optional

import torch.nn as nn


class Activation:
    """`Activation(name)()` -> nn.Module (None / "Identity" -> nn.Identity)."""

    def __init__(self, act_type: Optional[str]) -> None:
        self.type = act_type or "Identity"

    def __call__(self) -> nn.Module:
        if self.type in ("Identity", "None"):
            return nn.Identity()
        if not hasattr(nn, self.type):
            raise ValueError(f"unknown activation {self.type!r}")
        return getattr(nn, self.type)()
