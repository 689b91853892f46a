"""Tucker-2 decomposition of the k x k convolutions of a model (offline tool), same API as the reference's
scripts/tensor_decomposition/decomposition.py:

    decompose_model(model, loss_thr=0.1, prune_step=0.01)      in place               (:237-339)
    tucker_decomposition_conv_layer(conv) -> nn.Sequential     1x1 -> k x k -> 1x1    (:363-424)
    estimate_ranks(conv) -> [R0, R1]                           EVBMF on both unfoldings (:342-360)
    EVBMF(Y) -> (U, S, V, post)                                analytic empirical-VB MF (:81-206)

The reference runs this on the CPU with numpy + tensorly==0.6.0 (environment.yml:50, README.md:295). Here the linear
algebra (SVDs of the unfoldings, the HOOI sweeps of `partial_tucker`, the acceptance test convolutions) runs in torch
on whatever device the layer's weights live on -- on a B200 that is cuSOLVER / cuDNN through torch, which is fine for
an offline tool -- and only EVBMF's bounded scalar minimisation over the singular values (scipy) runs on the host.
The chains it emits are what `ayolov2_b200.engine` compiles into ONE fused kernel launch (csrc/conv_chain.cu).

tensorly's `partial_tucker(tensor, modes=[0, 1], rank, init="svd", n_iter_max=100, tol=1e-4)` is HOOI:
factors from the leading left singular vectors of each unfolding, then alternating updates
factor_i <- leading left singular vectors of unfold(tensor x_{j != i} factor_j^T, mode_i) until the relative
reconstruction error changes by less than tol. Singular vectors are unique up to sign; the chain is sign-invariant.
"""
from __future__ import annotations

import logging
from copy import deepcopy
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.utils.prune as prune
from torch import nn

LOGGER = logging.getLogger(__name__)


# ---------------------------------------------------------------------------------------------------------------
# EVBMF (Nakajima et al., "Global analytic solution of fully-observed variational Bayesian matrix factorization")
# ---------------------------------------------------------------------------------------------------------------
def tau(x: np.ndarray, alpha: float) -> np.ndarray:
    """decomposition.py:25-35."""
    return 0.5 * (x - (1 + alpha) + np.sqrt((x - (1 + alpha)) ** 2 - 4 * alpha))


def EVBsigma2(sigma2: float, L: int, M: int, s: np.ndarray, residual: float, xubar: float) -> float:
    """Free energy as a function of the noise variance (decomposition.py:38-78)."""
    H = len(s)
    alpha = L / M
    x = s ** 2 / (M * sigma2)
    z1 = x[x > xubar]
    z2 = x[x <= xubar]
    tau_z1 = tau(z1, alpha)
    term1 = np.sum(z2 - np.log(z2))
    term2 = np.sum(z1 - tau_z1)
    term3 = np.sum(np.log(np.divide(tau_z1 + 1, z1)))
    term4 = alpha * np.sum(np.log(tau_z1 / alpha + 1))
    return term1 + term2 + term3 + term4 + residual / (M * sigma2) + (L - H) * np.log(sigma2)


def _svd(Y: torch.Tensor):
    """Thin SVD in float64 on Y's device."""
    u, s, vh = torch.linalg.svd(Y.detach().double(), full_matrices=False)
    return u, s, vh


def EVBMF(Y: Union[torch.Tensor, np.ndarray], sigma2: Optional[float] = None, H: Optional[int] = None
          ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, Dict[str, np.ndarray]]:
    """Analytic EVBMF of Y (L x M, L <= M): returns (U[:, :pos], diag(d), V[:, :pos], post) like decomposition.py:81-206."""
    from scipy.optimize import minimize_scalar

    Yt = torch.as_tensor(Y)
    L, M = Yt.shape
    if H is None:
        H = L
    alpha = L / M
    tauubar = 2.5129 * np.sqrt(alpha)
    u, s_t, vh = _svd(Yt)
    U = u[:, :H].cpu().numpy()
    s = s_t[:H].cpu().numpy()
    V = vh[:H].T.cpu().numpy()
    residual = 0.0
    if H < L:
        residual = float(np.sum(np.sum(Yt.double().cpu().numpy() ** 2) - np.sum(s ** 2)))
    if sigma2 is None:
        xubar = (1 + tauubar) * (1 + alpha / tauubar)
        eH_ub = int(np.min([np.ceil(L / (1 + alpha)) - 1, H]))
        upper_bound = (np.sum(s ** 2) + residual) / (L * M)
        lower_bound = np.max([s[eH_ub] ** 2 / (M * xubar), np.mean(s[eH_ub:] ** 2) / M])
        sigma2 = minimize_scalar(EVBsigma2, args=(L, M, s, residual, xubar), bounds=[lower_bound, upper_bound],
                                 method="Bounded").x
    threshold = np.sqrt(M * sigma2 * (1 + tauubar) * (1 + alpha / tauubar))
    pos = int(np.sum(s > threshold))
    sp = s[:pos]
    d = np.multiply(sp / 2, 1 - np.divide((L + M) * sigma2, sp ** 2)
                    + np.sqrt((1 - np.divide((L + M) * sigma2, sp ** 2)) ** 2 - 4 * L * M * sigma2 ** 2 / sp ** 4))
    post: Dict[str, np.ndarray] = {k: np.zeros(H) for k in ("ma", "mb", "sa2", "sb2", "cacb")}
    tau_ = np.multiply(d, sp) / (M * sigma2)
    delta = np.multiply(np.sqrt(np.divide(M * d, L * sp)), 1 + alpha / tau_)
    post["ma"][:pos] = np.sqrt(np.multiply(d, delta))
    post["mb"][:pos] = np.sqrt(np.divide(d, delta))
    post["sa2"][:pos] = np.divide(sigma2 * delta, sp)
    post["sb2"][:pos] = np.divide(sigma2, np.multiply(delta, sp))
    post["cacb"][:pos] = np.sqrt(np.multiply(d, sp) / (L * M))
    post["sigma2"] = sigma2  # type: ignore
    post["F"] = 0.5 * (L * M * np.log(2 * np.pi * sigma2) + (residual + np.sum(s ** 2)) / sigma2
                       + np.sum(M * np.log(tau_ + 1) + L * np.log(tau_ / alpha + 1) - M * tau_))
    return U[:, :pos], np.diag(d), V[:, :pos], post


def _unfold(t: torch.Tensor, mode: int) -> torch.Tensor:
    return torch.moveaxis(t, mode, 0).reshape(t.shape[mode], -1)


def estimate_ranks(layer: nn.Conv2d) -> List[int]:
    """[rank of the mode-0 (out-channel) unfolding, rank of the mode-1 (in-channel) unfolding] (decomposition.py:342-360)."""
    w = layer.weight.data
    _, d0, _, _ = EVBMF(_unfold(w, 0))
    _, d1, _, _ = EVBMF(_unfold(w, 1))
    return [d0.shape[0], d1.shape[1]]


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
    """Conv2d(Cin, R1, 1) -> Conv2d(R1, R0, k, stride, pad) -> Conv2d(R0, Cout, 1, bias=orig) (decomposition.py:363-424)."""
    ranks = estimate_ranks(layer) if ranks is None else list(ranks)
    LOGGER.info("%s : VBMF Estimated ranks:  %s", layer, ranks)
    core, (last, first) = partial_tucker(layer.weight.data, ranks)
    dev, dt = layer.weight.device, layer.weight.dtype
    first_layer = nn.Conv2d(first.shape[0], first.shape[1], 1, 1, 0, dilation=layer.dilation, bias=False, device=dev, dtype=dt)
    core_layer = nn.Conv2d(core.shape[1], core.shape[0], layer.kernel_size, layer.stride, layer.padding, layer.dilation,
                           bias=False, device=dev, dtype=dt)
    last_layer = nn.Conv2d(last.shape[1], last.shape[0], 1, 1, 0, dilation=layer.dilation, bias=layer.bias is not None,
                           device=dev, dtype=dt)
    if layer.bias is not None:
        last_layer.bias.data = layer.bias.data
    first_layer.weight.data = torch.transpose(first, 1, 0).unsqueeze(-1).unsqueeze(-1).contiguous()
    last_layer.weight.data = last.unsqueeze(-1).unsqueeze(-1).contiguous()
    core_layer.weight.data = core.contiguous()
    return nn.Sequential(first_layer, core_layer, last_layer)


def decompose_layer_evaluation(layer: nn.Conv2d, test_input: torch.Tensor, origin_out: torch.Tensor
                               ) -> Tuple[Optional[nn.Sequential], Union[torch.Tensor, float]]:
    """decomposition.py:209-234: (chain, mean |difference| on the probe input) or (None, inf)."""
    try:
        chain = tucker_decomposition_conv_layer(deepcopy(layer))
    except ValueError:
        LOGGER.info("Decompose tensor failed.")
        return None, float("inf")
    with torch.no_grad():
        out = chain(test_input)
    return chain, torch.abs(origin_out - out).sum() / origin_out.numel()


def decompose_model(model: nn.Module, loss_thr: float = 0.1, prune_step: float = 0.01) -> None:
    """In place: every k x k (k > 1) Conv2d that is the `.conv` of its parent (or an entry of a ModuleList) is replaced by
    its Tucker-2 chain when the chain reproduces the layer on a random probe within `loss_thr`; a bisection over the
    L1-unstructured pruning ratio (step `prune_step`) then looks for the sparsest weights whose chain still passes
    (decomposition.py:237-339, including its RNG call order: one torch.rand probe per candidate layer)."""
    for i, (name, module) in enumerate(model.named_children()):
        if len(list(module.children())) > 0:
            decompose_model(module, loss_thr=loss_thr, prune_step=prune_step)
        if not isinstance(module, nn.Conv2d):
            continue
        conv = model[i] if isinstance(model, nn.ModuleList) else getattr(model, "conv", None)
        if conv is not module or conv.kernel_size == (1, 1):
            continue
        test_input = torch.rand((1024, *conv.weight.shape[1:])).to(conv.weight.device, conv.weight.dtype)
        with torch.no_grad():
            origin_out = conv(test_input)
        candidate, loss = decompose_layer_evaluation(conv, test_input, origin_out)
        LOGGER.info("%s (Prune: %.3f): Loss(mean): %s, ", name, 0.0, loss)
        chosen = candidate if loss < loss_thr else None
        search = loss < loss_thr and prune_step > 0
        lo, hi = 0.0, 1.0
        ratio = (lo + hi) / 2
        while search:
            pruned = deepcopy(conv)
            if ratio > 0.0:
                prune.l1_unstructured(pruned, name="weight", amount=ratio)
                prune.remove(pruned, "weight")
            candidate, loss = decompose_layer_evaluation(pruned, test_input, origin_out)
            LOGGER.info("%s (Prune: %.3f): Loss(mean): %s, ", name, ratio, loss)
            if loss < loss_thr:
                lo, chosen = ratio, candidate
            else:
                hi = ratio
            nxt = (lo + hi) / 2
            if abs(ratio - nxt) == 0 or abs(ratio - nxt) < prune_step:
                break
            ratio = nxt
        if chosen is None:
            LOGGER.info("    |---------- Skip switching to decomposed conv.")
            continue
        for attr in ("in_channels", "out_channels", "kernel_size"):
            setattr(chosen, attr, getattr(conv, attr))
        if isinstance(model, nn.ModuleList):
            model[i] = chosen
        else:
            model.conv = chosen
        LOGGER.info("    |---------- Switching conv to decomposed conv")
    if hasattr(model, "invalidate_engine"):
        model.invalidate_engine()
