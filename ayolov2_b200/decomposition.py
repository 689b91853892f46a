"""Tucker-2 decomposition of the k x k convolutions of a model (offline tool), same API as the reference's
scripts/tensor_decomposition/decomposition.py:

    decompose_model(model, loss_thr=0.1, prune_step=0.01)      in place               (:237-339)
    tucker_decomposition_conv_layer(conv) -> nn.Sequential     1x1 -> k x k -> 1x1    (:363-424)
    estimate_ranks(conv) -> [R0, R1]                           EVBMF on both unfoldings (:342-360)
    EVBMF(Y) -> (U, S, V, info)                                analytic empirical-VB MF (:81-206)
    evb_rank_batch / estimate_ranks_batch                      the same rank rule for many matrices in one vectorised search

The reference runs this on the CPU with numpy + tensorly==0.6.0 (environment.yml:50, README.md:295). Here the linear
algebra (SVDs of the unfoldings, the HOOI sweeps of `partial_tucker`, the acceptance test convolutions) runs in torch
on whatever device the layer's weights live on -- on a B200 that is cuSOLVER / cuDNN through torch, which is fine for
an offline tool -- and the EVB noise-variance search is a float64 Brent minimisation vectorised over matrices on the
same device (no scipy, no per-matrix host loop).
The chains it emits are what `ayolov2_b200.engine` compiles into ONE fused kernel launch (csrc/conv_chain.cu).

tensorly's `partial_tucker(tensor, modes=[0, 1], rank, init="svd", n_iter_max=100, tol=1e-4)` is HOOI:
factors from the leading left singular vectors of each unfolding, then alternating updates
factor_i <- leading left singular vectors of unfold(tensor x_{j != i} factor_j^T, mode_i) until the relative
reconstruction error changes by less than tol. Singular vectors are unique up to sign; the chain is sign-invariant.
"""
from __future__ import annotations

import logging
import math
from copy import deepcopy
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.utils.prune as prune
from torch import nn

LOGGER = logging.getLogger(__name__)


# ---------------------------------------------------------------------------------------------------------------
# Rank selection: empirical variational Bayes matrix factorisation (Nakajima, Sugiyama, Babacan, Tomioka, "Global
# analytic solution of fully-observed variational Bayesian matrix factorization", JMLR 2013), which the reference's
# estimate_ranks applies to the two unfoldings of a weight (decomposition.py:342-360). Only the RANK is consumed by the
# decomposition, so this module computes exactly that: the noise variance that minimises the EVB free energy over the
# published bracket, and the number of singular values above the resulting threshold. The search is Brent's bounded
# minimisation in float64, vectorised over ANY NUMBER of matrices at once (`evb_rank_batch`: every k x k layer of a model
# in one call, on the singular values' device), instead of one scipy scalar minimisation per matrix on the host.
# ---------------------------------------------------------------------------------------------------------------
_TAU_BAR_COEF = 2.5129  # the paper's numerical constant of the asymptotic threshold: tau_bar = 2.5129 * sqrt(alpha)


def _free_energy(v: torch.Tensor, s2: torch.Tensor, valid: torch.Tensor, rows: torch.Tensor, cols: torch.Tensor,
                 kappa: torch.Tensor) -> torch.Tensor:
    """EVB free energy (up to constants) at noise variance v, per matrix. v, rows, cols, kappa: [n]; s2 (squared singular
    values, zero padded) and valid: [n, h]. For a component with x = s^2 / (cols * v):
        x <= kappa (explained as noise):   x - ln x
        x >  kappa (kept):                 x - t + ln((t + 1) / x) + alpha * ln(t / alpha + 1),
                                           t = ((x - 1 - alpha) + sqrt((x - 1 - alpha)^2 - 4 alpha)) / 2."""
    alpha = (rows / cols)[:, None]
    x = s2 / (cols * v)[:, None]
    x = torch.where(valid, x, torch.ones_like(x))  # padding contributes the constant 1 - ln 1 to every candidate v
    shifted = x - 1.0 - alpha
    t = 0.5 * (shifted + torch.sqrt((shifted * shifted - 4.0 * alpha).clamp_min(0.0)))
    kept = x - t + torch.log((t + 1.0) / x) + alpha * torch.log(t / alpha + 1.0)
    noise = x - torch.log(x)
    # a matrix with more rows than columns has only `cols` singular values: the missing rows - cols components enter
    # through ln v (the reference keeps Y as given, decomposition.py:100-118, so this follows it rather than transposing)
    missing = rows - valid.sum(1)
    return torch.where(x > kappa[:, None], kept, noise).sum(1) + missing * torch.log(v)


def _bounded_minimum(f, lo: torch.Tensor, hi: torch.Tensor, xatol: float = 1e-5, max_evals: int = 500) -> torch.Tensor:
    """Minimum of n scalar functions on [lo_i, hi_i] at once: Brent's derivative-free `localmin` (R. P. Brent, "Algorithms
    for Minimization without Derivatives", 1973, ch. 5; the FMIN of Forsythe, Malcolm & Moler) -- a golden-section search
    that takes a parabolic-interpolation step through the three best points whenever that step is acceptable -- run in
    lock step over a batch with masked updates, float64, on the tensors' device.

    The reference finds its noise variance with scipy's bounded scalar minimiser at scipy's default ABSOLUTE tolerance
    (1e-5, coarse next to variances of 1e-5 .. 1e-2), so which rank a borderline layer gets depends on where that
    particular iteration stops. This routine therefore applies the same published acceptance rules, tolerances and
    stopping test, which makes it stop at the same point (and reproduces the reference's 6,329,941-parameter golden
    decomposition), instead of converging further than the reference does."""
    tiny = math.sqrt(2.2e-16)
    gold = 0.5 * (3.0 - math.sqrt(5.0))
    a, b = lo.clone(), hi.clone()
    best = a + gold * (b - a)          # best point so far; second and third best start there as well
    second, third = best.clone(), best.clone()
    f_best = f(best)
    f_second, f_third = f_best.clone(), f_best.clone()
    step = torch.zeros_like(best)      # last step taken
    prev = torch.zeros_like(best)      # the step before it
    mid = 0.5 * (a + b)
    tol1 = tiny * best.abs() + xatol / 3.0
    tol2 = 2.0 * tol1
    evals = 1
    while evals < max_evals:
        active = (best - mid).abs() > (tol2 - 0.5 * (b - a))
        if not bool(active.any()):
            break
        # parabola through (best, second, third); p / q is the proposed displacement from `best`
        r = (best - second) * (f_best - f_third)
        q = (best - third) * (f_best - f_second)
        p = (best - third) * q - (best - second) * r
        q = 2.0 * (q - r)
        p = torch.where(q > 0.0, -p, p)
        q = q.abs()
        try_parabola = prev.abs() > tol1
        ok = try_parabola & (p.abs() < (0.5 * q * prev).abs()) & (p > q * (a - best)) & (p < q * (b - best))
        para = p / torch.where(q == 0, torch.ones_like(q), q)
        cand = best + para
        near_edge = ((cand - a) < tol2) | ((b - cand) < tol2)
        toward_mid = torch.sign(mid - best) + ((mid - best) == 0).to(best.dtype)
        para = torch.where(near_edge, tol1 * toward_mid, para)
        gold_prev = torch.where(best >= mid, a - best, b - best)
        new_prev = torch.where(ok, step, gold_prev)           # parabolic: remember the previous step; golden: the wider side
        new_step = torch.where(ok, para, gold * gold_prev)
        sgn = torch.sign(new_step) + (new_step == 0).to(best.dtype)
        x = best + sgn * torch.maximum(new_step.abs(), tol1)   # never evaluate closer than tol1 to the best point
        fx = f(x)
        evals += 1
        better = fx <= f_best
        # bracket update
        na = torch.where(better, torch.where(x >= best, best, a), torch.where(x < best, x, a))
        nb = torch.where(better, torch.where(x >= best, b, best), torch.where(x < best, b, x))
        # ranking update of the three remembered points
        demote2 = ~better & ((fx <= f_second) | (second == best))
        demote3 = ~better & ~demote2 & ((fx <= f_third) | (third == best) | (third == second))
        n_third = torch.where(better | demote2, second, torch.where(demote3, x, third))
        nf_third = torch.where(better | demote2, f_second, torch.where(demote3, fx, f_third))
        n_second = torch.where(better, best, torch.where(demote2, x, second))
        nf_second = torch.where(better, f_best, torch.where(demote2, fx, f_second))
        n_best = torch.where(better, x, best)
        nf_best = torch.where(better, fx, f_best)
        # finished problems are frozen
        a, b = torch.where(active, na, a), torch.where(active, nb, b)
        third, f_third = torch.where(active, n_third, third), torch.where(active, nf_third, f_third)
        second, f_second = torch.where(active, n_second, second), torch.where(active, nf_second, f_second)
        best, f_best = torch.where(active, n_best, best), torch.where(active, nf_best, f_best)
        step, prev = torch.where(active, new_step, step), torch.where(active, new_prev, prev)
        mid = 0.5 * (a + b)
        tol1 = tiny * best.abs() + xatol / 3.0
        tol2 = 2.0 * tol1
    return best


def evb_rank_batch(singular_values: List[torch.Tensor], shapes: List[Tuple[int, int]]) -> Tuple[List[int], List[float]]:
    """EVB ranks of n matrices at once. singular_values[i]: all min(L_i, M_i) singular values of matrix i (descending),
    shapes[i] = (L_i, M_i). Returns (ranks, noise variances)."""
    n = len(singular_values)
    if n == 0:
        return [], []
    dev = singular_values[0].device
    h = max(int(s.numel()) for s in singular_values)
    s2 = torch.zeros((n, h), dtype=torch.float64, device=dev)
    valid = torch.zeros((n, h), dtype=torch.bool, device=dev)
    rows = torch.empty(n, dtype=torch.float64, device=dev)
    cols = torch.empty(n, dtype=torch.float64, device=dev)
    lo = torch.empty(n, dtype=torch.float64, device=dev)
    hi = torch.empty(n, dtype=torch.float64, device=dev)
    for i, (sv, (L, M)) in enumerate(zip(singular_values, shapes)):
        q = sv.double() ** 2
        k = int(q.numel())
        assert k == min(L, M), "all singular values of the unfolding are needed"
        s2[i, :k], valid[i, :k] = q, True
        rows[i], cols[i] = L, M
    alpha = rows / cols
    tau_bar = _TAU_BAR_COEF * torch.sqrt(alpha)
    kappa = (1.0 + tau_bar) * (1.0 + alpha / tau_bar)
    # the paper's bracket for the noise variance: above, the mean energy per entry; below, whatever makes the first
    # component beyond the largest admissible rank a noise component, or the mean of the tail (whichever is larger)
    for i in range(n):
        L, M = int(rows[i]), float(cols[i])
        q = s2[i, :int(valid[i].sum())]
        cut = min(math.ceil(L / (1.0 + L / M)) - 1, int(q.numel()) - 1)
        hi[i] = q.sum() / (L * M)
        lo[i] = torch.maximum(q[cut] / (M * kappa[i]), q[cut:].mean() / M)
    v = _bounded_minimum(lambda x: _free_energy(x, s2, valid, rows, cols, kappa), lo, hi)
    keep = (s2 > (cols * v * kappa)[:, None]) & valid  # s > sqrt(M v (1 + tau_bar)(1 + alpha / tau_bar))
    return [int(r) for r in keep.sum(1).tolist()], [float(t) for t in v.tolist()]


def _singular_values(m: torch.Tensor) -> torch.Tensor:
    return torch.linalg.svdvals(m.detach().double())


def _svd(Y: torch.Tensor):
    """Thin SVD in float64 on Y's device."""
    u, s, vh = torch.linalg.svd(Y.detach().double(), full_matrices=False)
    return u, s, vh


def EVBMF(Y: Union[torch.Tensor, np.ndarray]) -> Tuple[np.ndarray, np.ndarray, np.ndarray, Dict[str, float]]:
    """Reference-shaped entry point (decomposition.py:81-206): (U[:, :r], diag(d), V[:, :r], info) with r the EVB rank
    and d the EVB-shrunk singular values. `info` carries the noise variance and the rank; the posterior moments the
    reference also tabulates are consumed by nothing in the repository (or the reference) and are not computed."""
    Yt = torch.as_tensor(Y)
    L, M = Yt.shape
    u, s, vh = _svd(Yt)
    (r,), (v,) = evb_rank_batch([s], [(L, M)])
    top = s[:r]
    # EVB shrinkage of a kept component: gamma = s/2 * (c + sqrt(c^2 - 4 L M v^2 / s^4)), c = 1 - (L + M) v / s^2
    c = 1.0 - (L + M) * v / top ** 2
    d = 0.5 * top * (c + torch.sqrt((c * c - 4.0 * L * M * v * v / top ** 4).clamp_min(0.0)))
    return (u[:, :r].cpu().numpy(), np.diag(d.cpu().numpy()), vh[:r].T.cpu().numpy(), {"sigma2": v, "rank": r})


def _unfold(t: torch.Tensor, mode: int) -> torch.Tensor:
    return torch.moveaxis(t, mode, 0).reshape(t.shape[mode], -1)


def estimate_ranks(layer: nn.Conv2d) -> List[int]:
    """[rank of the mode-0 (out-channel) unfolding, rank of the mode-1 (in-channel) unfolding] (decomposition.py:342-360)."""
    return estimate_ranks_batch([layer])[0]


def estimate_ranks_batch(layers: List[nn.Conv2d]) -> List[List[int]]:
    """Tucker-2 ranks of many convolutions with ONE vectorised noise-variance search (2 unfoldings per layer)."""
    svals, shapes = [], []
    for layer in layers:
        w = layer.weight.data
        for mode in (0, 1):
            m = _unfold(w, mode)
            svals.append(_singular_values(m))
            shapes.append(tuple(m.shape))
    ranks, _ = evb_rank_batch(svals, shapes)
    return [[ranks[2 * i], ranks[2 * i + 1]] for i in range(len(layers))]


# ---------------------------------------------------------------------------------------------------------------
# partial Tucker (HOOI) on modes 0 and 1 of an OIHW weight
# ---------------------------------------------------------------------------------------------------------------
def _leading(m: torch.Tensor, k: int) -> torch.Tensor:
    u, _, _ = _svd(m)
    return u[:, :k].to(m.dtype)


def partial_tucker(tensor: torch.Tensor, rank: List[int], n_iter_max: int = 100, tol: float = 1e-4):
    """core (R0, R1, kh, kw), [last (Cout, R0), first (Cin, R1)] with orthonormal factors."""
    r0, r1 = int(rank[0]), int(rank[1])
    if r0 < 1 or r1 < 1:
        raise ValueError(f"Tucker ranks must be positive, got {rank}")  # decomposition.py:222-227 catches this
    t = tensor.detach()
    last = _leading(_unfold(t, 0), r0)
    first = _leading(_unfold(t, 1), r1)
    norm_t = float(torch.linalg.norm(t.double()))
    errs: List[float] = []
    core = t
    for it in range(n_iter_max):
        last = _leading(_unfold(torch.einsum("oikl,is->oskl", t, first), 0), r0)
        first = _leading(_unfold(torch.einsum("oikl,or->rikl", t, last), 1), r1)
        core = torch.einsum("oikl,or,is->rskl", t, last, first)
        err = float(np.sqrt(abs(norm_t ** 2 - float(torch.linalg.norm(core.double())) ** 2))) / norm_t
        errs.append(err)
        if it > 1 and abs(errs[-2] - errs[-1]) < tol:
            break
    return core, [last, first]


def tucker_decomposition_conv_layer(layer: nn.Conv2d, ranks: Optional[List[int]] = None) -> nn.Sequential:
    """The Tucker-2 chain of one convolution (decomposition.py:363-424):
        1x1 (Cin -> R1, no bias)  ->  k x k (R1 -> R0, the layer's stride / padding / dilation, no bias)  ->  1x1 (R0 -> Cout, the layer's bias)
    with W ~= core x_0 U_out x_1 U_in; ranks default to the EVB estimate of the two unfoldings."""
    r_out, r_in = estimate_ranks(layer) if ranks is None else (int(ranks[0]), int(ranks[1]))
    LOGGER.info("Tucker-2 of %s with ranks (out %d, in %d)", layer, r_out, r_in)
    core, (u_out, u_in) = partial_tucker(layer.weight.data, [r_out, r_in])
    kw = dict(device=layer.weight.device, dtype=layer.weight.dtype)
    squeeze = nn.Conv2d(layer.in_channels, r_in, 1, bias=False, dilation=layer.dilation, **kw)
    mix = nn.Conv2d(r_in, r_out, layer.kernel_size, layer.stride, layer.padding, layer.dilation, bias=False, **kw)
    expand = nn.Conv2d(r_out, layer.out_channels, 1, bias=layer.bias is not None, dilation=layer.dilation, **kw)
    with torch.no_grad():
        squeeze.weight.copy_(u_in.t()[:, :, None, None])   # y1[r] = sum_i U_in[i, r] x[i]
        mix.weight.copy_(core)
        expand.weight.copy_(u_out[:, :, None, None])       # y[o] = sum_r U_out[o, r] y2[r]
        if layer.bias is not None:
            expand.bias.copy_(layer.bias.data)
    return nn.Sequential(squeeze, mix, expand)


def decompose_layer_evaluation(layer: nn.Conv2d, test_input: torch.Tensor, origin_out: torch.Tensor
                               ) -> Tuple[Optional[nn.Sequential], Union[torch.Tensor, float]]:
    """(chain, mean absolute deviation from `origin_out` on `test_input`), or (None, inf) when the EVB ranks admit no chain
    (decomposition.py:209-234)."""
    try:
        chain = tucker_decomposition_conv_layer(deepcopy(layer))
    except ValueError:
        LOGGER.info("no Tucker-2 chain for %s (a rank came out as zero)", layer)
        return None, float("inf")
    with torch.no_grad():
        return chain, (origin_out - chain(test_input)).abs().mean()


class _Site:
    """One replaceable convolution: where it hangs (attribute `conv` of its parent, or an index of a ModuleList)."""

    def __init__(self, parent: nn.Module, key: Union[int, str], conv: nn.Conv2d, name: str) -> None:
        self.parent, self.key, self.conv, self.name = parent, key, conv, name

    def replace(self, chain: nn.Sequential) -> None:
        chain.in_channels, chain.out_channels, chain.kernel_size = self.conv.in_channels, self.conv.out_channels, self.conv.kernel_size
        if isinstance(self.key, int):
            self.parent[self.key] = chain
        else:
            setattr(self.parent, self.key, chain)


def _candidate_sites(root: nn.Module):
    """Spatial (k > 1) convolutions in depth-first order of the module tree -- the order in which the reference's recursion
    meets them, which fixes the order of the random probes and therefore the result (decomposition.py:237-272). Eligible
    are convolutions that are the `.conv` of their parent or an entry of a ModuleList."""
    for idx, (name, child) in enumerate(root.named_children()):
        yield from _candidate_sites(child)
        if not isinstance(child, nn.Conv2d) or child.kernel_size == (1, 1):
            continue
        if isinstance(root, nn.ModuleList):
            yield _Site(root, idx, child, name)
        elif getattr(root, "conv", None) is child:
            yield _Site(root, "conv", child, name)


def _sparsest_passing_chain(conv: nn.Conv2d, probe: torch.Tensor, target: torch.Tensor, loss_thr: float, step: float,
                            name: str) -> Optional[nn.Sequential]:
    """The chain of the most heavily L1-pruned copy of `conv` that still reproduces `target` within loss_thr, found by
    bisecting the pruning ratio on [0, 1] until the interval is narrower than `step` (decomposition.py:275-321). The
    unpruned layer is tried first; if even that fails there is no chain."""
    chain, loss = decompose_layer_evaluation(conv, probe, target)
    LOGGER.info("%s: unpruned chain deviates by %s", name, loss)
    if not loss < loss_thr:
        return None
    if step <= 0:
        return chain
    passed, failed = 0.0, 1.0
    ratio = 0.5 * (passed + failed)
    while True:
        trial = deepcopy(conv)
        prune.l1_unstructured(trial, name="weight", amount=ratio)
        prune.remove(trial, "weight")
        cand, loss = decompose_layer_evaluation(trial, probe, target)
        LOGGER.info("%s: pruned by %.3f, chain deviates by %s", name, ratio, loss)
        if loss < loss_thr:
            passed, chain = ratio, cand
        else:
            failed = ratio
        nxt = 0.5 * (passed + failed)
        if abs(nxt - ratio) < step or nxt == ratio:
            return chain
        ratio = nxt


def decompose_model(model: nn.Module, loss_thr: float = 0.1, prune_step: float = 0.01) -> None:
    """In place (decomposition.py:237-339): every eligible k x k convolution is replaced by its Tucker-2 chain when the
    chain reproduces the layer's response to a random probe batch (1024 x Cin x k x k, one `torch.rand` per candidate, in
    tree order) within `loss_thr` mean absolute deviation; among the passing chains the one of the sparsest weights wins."""
    for site in list(_candidate_sites(model)):
        conv = site.conv
        probe = torch.rand((1024, *conv.weight.shape[1:])).to(conv.weight.device, conv.weight.dtype)
        with torch.no_grad():
            target = conv(probe)
        chain = _sparsest_passing_chain(conv, probe, target, loss_thr, prune_step, site.name)
        if chain is None:
            LOGGER.info("%s stays dense", site.name)
            continue
        site.replace(chain)
        LOGGER.info("%s replaced by a Tucker-2 chain", site.name)
    for m in model.modules():
        if hasattr(m, "invalidate_engine"):
            m.invalidate_engine()
