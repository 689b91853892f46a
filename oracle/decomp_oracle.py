"""CPU oracle for the offline Tucker-2 factorisation. TEST INFRASTRUCTURE ONLY.

Restates, in numpy float64 / torch on the CPU:
  * EVBMF rank estimation              scripts/tensor_decomposition/decomposition.py:25-206, 342-360
  * tensorly==0.6.0 `tl.base.unfold` and `tensorly.decomposition.partial_tucker(modes=[0,1], init="svd",
    n_iter_max=100, tol=1e-4)` (HOOI). tensorly is a third-party dependency pinned in the reference's
    environment.yml:50, absent from /root/reference and not installed here; its published algorithm is:
        factors[i] <- leading rank[i] left singular vectors of unfold(tensor, mode_i)          (SVD init)
        repeat: for each mode i: factors[i] <- leading left singular vectors of
                    unfold(tensor x_{j != i} factors[j]^T, mode_i);
                core <- tensor x_j factors[j]^T;  err <- sqrt(|T|^2 - |core|^2) / |T|;
                stop when iteration > 1 and |err[-2] - err[-1]| < tol.
  * the three-conv chain construction  decomposition.py:363-424
  * the acceptance / prune-ratio bisection loop of `decompose_model`  decomposition.py:209-339

PIN: `load_reference()` imports the UNMODIFIED reference module with a `tensorly` stub whose `unfold` /
`partial_tucker` are the restatements above; running the reference's own test flow
(tests/test_tensor_decomposition.py:22-49) on its fixture checkpoint then reproduces the reference's golden
numbers -- 6,329,941 parameters after decomposition and a full-forward loss < 0.015 -- which pins both the HOOI
restatement and (through this file's own `decompose_model`, compared layer by layer) the rest.
Singular vectors are unique up to sign; the chain's product first x core x last is sign-invariant, so parity
is asserted on ranks, parameter counts and chain outputs, not on individual factor entries.
"""
from __future__ import annotations

import os
import sys
import types
from copy import deepcopy
from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import ref_import


# ---------------------------------------------------------------------------------------------------------------
# tensorly 0.6.0 restatement
# ---------------------------------------------------------------------------------------------------------------
def unfold(t: torch.Tensor, mode: int) -> torch.Tensor:
    """tl.base.unfold: mode-`mode` fibres become the rows' index, remaining modes keep their order (C order)."""
    return torch.moveaxis(t, mode, 0).reshape(t.shape[mode], -1)


def _mode_dot_t(t: torch.Tensor, factor: torch.Tensor, mode: int) -> torch.Tensor:
    """tensor x_mode factor^T  (factor: (dim, rank))  -> mode dimension becomes `rank`."""
    moved = torch.moveaxis(t, mode, -1)
    return torch.moveaxis(moved @ factor, -1, mode)


def _leading_left_vectors(m: torch.Tensor, k: int) -> torch.Tensor:
    u, _, _ = torch.linalg.svd(m.double(), full_matrices=False)
    return u[:, :k].to(m.dtype)


def partial_tucker(tensor: torch.Tensor, modes: List[int], rank: List[int], n_iter_max: int = 100, init: str = "svd",
                   tol: float = 10e-5, **_):
    """tensorly.decomposition.partial_tucker (0.6.0) for init='svd', no mask. Returns (core, factors)."""
    assert init == "svd"
    rank = tuple(int(r) for r in rank)
    factors = [_leading_left_vectors(unfold(tensor, m), rank[i]) for i, m in enumerate(modes)]
    norm_tensor = float(torch.linalg.norm(tensor.double()))
    rec_errors: List[float] = []
    core = tensor
    for iteration in range(n_iter_max):
        for i, mode in enumerate(modes):
            approx = tensor
            for j, mj in enumerate(modes):
                if j != i:
                    approx = _mode_dot_t(approx, factors[j], mj)
            factors[i] = _leading_left_vectors(unfold(approx, mode), rank[i])
        core = tensor
        for j, mj in enumerate(modes):
            core = _mode_dot_t(core, factors[j], mj)
        rec = float(np.sqrt(abs(norm_tensor ** 2 - float(torch.linalg.norm(core.double())) ** 2))) / norm_tensor
        rec_errors.append(rec)
        if iteration > 1 and tol and abs(rec_errors[-2] - rec_errors[-1]) < tol:
            break
    return core, factors


# ---------------------------------------------------------------------------------------------------------------
# EVBMF (decomposition.py:25-206) restated on the singular values only (the reference discards U, V and `post`)
# ---------------------------------------------------------------------------------------------------------------
def _tau(x: np.ndarray, alpha: float) -> np.ndarray:
    return 0.5 * (x - (1 + alpha) + np.sqrt((x - (1 + alpha)) ** 2 - 4 * alpha))  # :25-35


def _evb_sigma2(sigma2: float, L: int, M: int, s: np.ndarray, residual: float, xubar: float) -> float:
    H = len(s)  # :38-78
    alpha = L / M
    x = s ** 2 / (M * sigma2)
    z1 = x[x > xubar]
    z2 = x[x <= xubar]
    tz1 = _tau(z1, alpha)
    return (np.sum(z2 - np.log(z2)) + np.sum(z1 - tz1) + np.sum(np.log(np.divide(tz1 + 1, z1)))
            + alpha * np.sum(np.log(tz1 / alpha + 1)) + residual / (M * sigma2) + (L - H) * np.log(sigma2))


def evbmf_rank(Y: np.ndarray) -> int:
    """Number of singular values EVBMF keeps (= `diag.shape[0]` at decomposition.py:356-359), sigma2 estimated."""
    return evbmf_rank_sigma2(Y)[0]


def evbmf_rank_sigma2(Y: np.ndarray):
    """(rank, estimated noise variance) of decomposition.py:81-206 with sigma2 = None, H = None."""
    from scipy.optimize import minimize_scalar

    L, M = Y.shape
    H = L
    alpha = L / M
    tauubar = 2.5129 * np.sqrt(alpha)
    s = np.linalg.svd(np.asarray(Y), compute_uv=False)[:H]
    residual = 0.0
    xubar = (1 + tauubar) * (1 + alpha / tauubar)
    eH_ub = int(np.min([np.ceil(L / (1 + alpha)) - 1, H]))
    upper = (np.sum(s ** 2) + residual) / (L * M)
    lower = np.max([s[eH_ub] ** 2 / (M * xubar), np.mean(s[eH_ub:] ** 2) / M])
    opt = minimize_scalar(_evb_sigma2, args=(L, M, s, residual, xubar), bounds=[lower, upper], method="Bounded")
    sigma2 = opt.x
    threshold = np.sqrt(M * sigma2 * (1 + tauubar) * (1 + alpha / tauubar))
    return int(np.sum(s > threshold)), float(sigma2)


def estimate_ranks(weight: torch.Tensor) -> List[int]:
    """decomposition.py:342-360: [rank of the mode-0 unfolding, rank of the mode-1 unfolding]."""
    return [evbmf_rank(unfold(weight, 0).numpy()), evbmf_rank(unfold(weight, 1).numpy())]


def tucker_chain(layer: nn.Conv2d, ranks: Optional[List[int]] = None) -> nn.Sequential:
    """decomposition.py:363-424. Raises ValueError when a rank is 0 (tensorly does, :222-227 catches it)."""
    w = layer.weight.data
    ranks = estimate_ranks(w) if ranks is None else ranks
    if min(ranks) < 1:
        raise ValueError("rank 0")
    core, (last, first) = partial_tucker(w, modes=[0, 1], rank=ranks, init="svd")
    f = nn.Conv2d(first.shape[0], first.shape[1], 1, 1, 0, dilation=layer.dilation, bias=False)
    c = nn.Conv2d(core.shape[1], core.shape[0], layer.kernel_size, layer.stride, layer.padding, layer.dilation, bias=False)
    l = nn.Conv2d(last.shape[1], last.shape[0], 1, 1, 0, dilation=layer.dilation, bias=layer.bias is not None)
    if layer.bias is not None:
        l.bias.data = layer.bias.data
    f.weight.data = torch.transpose(first, 1, 0).unsqueeze(-1).unsqueeze(-1)
    l.weight.data = last.unsqueeze(-1).unsqueeze(-1)
    c.weight.data = core
    return nn.Sequential(f, c, l)


# ---------------------------------------------------------------------------------------------------------------
# The unmodified reference module, with the tensorly stub
# ---------------------------------------------------------------------------------------------------------------
_ref_mod = None


def load_reference():
    """Import scripts/tensor_decomposition/decomposition.py from /root/reference (build container only)."""
    global _ref_mod
    if _ref_mod is not None:
        return _ref_mod
    if not ref_import.available():
        raise RuntimeError("reference tree not available")
    ref_import.load()  # sys.path + matplotlib / kindle stubs
    tl = types.ModuleType("tensorly")
    tl.__dict__["__ay2_stub__"] = True
    tl.set_backend = lambda name: None
    tl.base = types.ModuleType("tensorly.base")
    tl.base.unfold = unfold
    tl.decomposition = types.ModuleType("tensorly.decomposition")
    tl.decomposition.partial_tucker = partial_tucker
    sys.modules.setdefault("tensorly", tl)
    sys.modules.setdefault("tensorly.base", tl.base)
    sys.modules.setdefault("tensorly.decomposition", tl.decomposition)
    from scripts.tensor_decomposition import decomposition  # type: ignore

    _ref_mod = decomposition
    return decomposition


def fixture_checkpoint_path() -> str:
    return os.path.join(ref_import.REF_ROOT, "tests", "res", "weights", "yolov5s_kindle.pt")
