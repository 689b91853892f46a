"""Shared helpers of the GPU parity tests: error metrics and a JSON record of every MEASURED error.

Each parity test calls `record(case, **numbers)`; the records are merged into `$AY2_PARITY_JSON` (default
`gpurun_out/r02_parity.json`, which gpurun brings back from the GPU box). The committed copy lives in
`profiles/r02_parity.json`: the asserted tolerances are the north-star ones, the file shows how far below them the
CUDA path actually is.
"""
from __future__ import annotations

import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _path() -> str:
    return os.environ.get("AY2_PARITY_JSON", os.path.join(ROOT, "gpurun_out", "r02_parity.json"))


def max_norm(got: torch.Tensor, ref: torch.Tensor) -> float:
    """max |got - ref| / max |ref|"""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


def rel_l2(got: torch.Tensor, ref: torch.Tensor) -> float:
    """||got - ref||_2 / ||ref||_2"""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-12))


def errs(got: torch.Tensor, ref: torch.Tensor) -> dict:
    return {"max_norm": max_norm(got, ref), "rel_l2": rel_l2(got, ref)}


def record(case: str, **numbers) -> None:
    path = _path()
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[case] = {k: (float(v) if isinstance(v, (int, float)) else v) for k, v in numbers.items()}
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:  # a read-only checkout must not fail a parity test
        pass
    print(f"[parity] {case}: " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in numbers.items()))
