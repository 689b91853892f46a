"""kindle.modules.concat — Concat(dimension). In the engine a Concat is pure addressing:
producers write into channel slices of one NHWC buffer."""
from __future__ import annotations

from typing import List, Union

import torch
import torch.nn as nn


class Concat(nn.Module):
    def __init__(self, dimension: int = 1) -> None:
        super().__init__()
        self.dimension = dimension

    def forward(self, x: Union[torch.Tensor, List[torch.Tensor]]) -> torch.Tensor:
        return torch.cat(list(x), dim=self.dimension)
