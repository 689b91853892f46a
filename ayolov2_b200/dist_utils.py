"""Multi-GPU plumbing (one process per GPU, torch.distributed): rank discovery, barrier, max-over-ranks timing, and
the flat-buffer gradient all-reduce of the training step.

The reference's only parallelism is data parallel: inference + NMS shards over images with no
exchange step (replicas), training exchanges gradients once per step through DDP (scripts/train/train_model_builder.py:
75-78,112-114 `dist.init_process_group("nccl" | "gloo")`, `DDP(model, device_ids=[LOCAL_RANK])`).
Backend is NCCL on GPUs; the same helpers run on gloo for the CPU tests (tests/test_dist_cpu.py)."""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def env_ranks() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torch.distributed.run environment (train.py:22-26)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: Optional[str] = None, device: Optional[torch.device] = None) -> int:
    """Initialise the default process group if WORLD_SIZE > 1. Returns the world size."""
    _, _, world = env_ranks()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() and dist.is_nccl_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return world


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Every multi-GPU time is reported as the max over ranks of the device-side measurement."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_batch(global_batch: int, rank: int, world: int) -> int:
    """Per-rank batch exactly as the reference splits it: cfg.batch_size // WORLD_SIZE (data_loader_utils.py:67)."""
    return global_batch // world


def allreduce_mean_(grads: Iterable[torch.Tensor], bucket_bytes: int = 256 << 20) -> int:
    """In-place mean all-reduce of a list of gradient tensors through flat buckets (DDP semantics: gradients are
    averaged over ranks, which is why the trainer multiplies the loss by world_size, yolo_trainer.py:325-326).
    One NCCL call per bucket; yolov5s (29 MB fp32) is a single message on NVLink/NVSwitch. Returns the bucket count."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    grads = [g for g in grads if g is not None]
    buckets: List[List[torch.Tensor]] = [[]]
    size = 0
    for g in grads:
        nb = g.numel() * g.element_size()
        if buckets[-1] and (size + nb > bucket_bytes or g.dtype != buckets[-1][0].dtype):
            buckets.append([])
            size = 0
        buckets[-1].append(g)
        size += nb
    n = 0
    for b in buckets:
        if not b:
            continue
        flat = torch.cat([g.reshape(-1) for g in b])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in b:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n += 1
    return n
