"""Tucker-2 convolution factorisation with FIXED ranks (offline, CPU), producing exactly the module structure the
reference's decomposition emits (scripts/tensor_decomposition/decomposition.py:363-424):

    Conv2d(Cin, R1, 1, bias=False) -> Conv2d(R1, R0, k, stride, pad, bias=False) -> Conv2d(R0, Cout, 1, bias=orig)

The reference picks data-dependent ranks with EVBMF and refines the factors with tensorly's HOOI (`partial_tucker`,
tensorly==0.6.0, not installed here); that offline search is out of scope this round. For the
runtime path (BASELINE.json configs[3]) what matters is the three-conv chain; this helper builds it with a truncated
HOSVD (= HOOI's SVD initialisation) at ranks `ceil(ratio * C)` so that benchmarks and parity tests are reproducible.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn


def tucker2_conv(conv: nn.Conv2d, rank_out: int, rank_in: int) -> nn.Sequential:
    """Truncated-HOSVD Tucker-2 of an OIHW weight along modes 0 (out) and 1 (in)."""
    w = conv.weight.detach().float().cpu()
    cout, cin, kh, kw = w.shape
    u_out, _, _ = torch.linalg.svd(w.reshape(cout, -1), full_matrices=False)
    u_in, _, _ = torch.linalg.svd(w.permute(1, 0, 2, 3).reshape(cin, -1), full_matrices=False)
    last = u_out[:, :rank_out]   # (Cout, R0)
    first = u_in[:, :rank_in]    # (Cin, R1)
    core = torch.einsum("oikl,or,is->rskl", w, last, first)  # (R0, R1, kh, kw)
    f = nn.Conv2d(cin, rank_in, 1, 1, 0, bias=False)
    c = nn.Conv2d(rank_in, rank_out, conv.kernel_size, conv.stride, conv.padding, conv.dilation, bias=False)
    l = nn.Conv2d(rank_out, cout, 1, 1, 0, bias=conv.bias is not None)
    with torch.no_grad():
        f.weight.copy_(first.t()[:, :, None, None])   # decomposition.py:419
        c.weight.copy_(core)                          # :420
        l.weight.copy_(last[:, :, None, None])        # :421
        if conv.bias is not None:
            l.bias.copy_(conv.bias.detach().float().cpu())
    return nn.Sequential(f, c, l).to(conv.weight.device)


def decompose_model_fixed(model: nn.Module, ratio: float = 0.5, skip_first: bool = True) -> List[str]:
    """In-place: every kindle Conv whose `.conv` is an nn.Conv2d with kernel != 1x1 (decomposition.py:271-272) gets the
    three-conv chain at ranks ceil(ratio*C). Returns the names of the replaced modules. The image-fed stem is skipped
    (`skip_first`): the engine compiles it through the space-to-depth path, which has no factored form yet."""
    replaced = []
    first_conv = next((m for m in model.modules() if hasattr(m, "conv") and isinstance(getattr(m, "conv"), nn.Conv2d)), None)
    for name, m in model.named_modules():
        conv = getattr(m, "conv", None)
        if type(m).__name__ not in ("Conv", "Focus") or not isinstance(conv, nn.Conv2d):
            continue
        if conv.kernel_size == (1, 1) or (skip_first and m is first_conv):
            continue
        r0 = max(1, math.ceil(ratio * conv.out_channels))
        r1 = max(1, math.ceil(ratio * conv.in_channels))
        m.conv = tucker2_conv(conv, r0, r1)
        m.in_channels, m.out_channels, m.kernel_size = conv.in_channels, conv.out_channels, conv.kernel_size  # :325-335
        replaced.append(name)
    if hasattr(model, "invalidate_engine"):
        model.invalidate_engine()
    return replaced
