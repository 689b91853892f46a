#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[2]): yolov5s fwd (batch-statistics BN) + ComputeLoss + backward +
SGD-nesterov/EMA step on synthetic 640x640 images and synthetic targets, one process per GPU (DDP-style gradient
mean all-reduce over NCCL when WORLD_SIZE > 1).

  python tools/bench_train.py [--batch 128] [--steps 10] [--warmup 3] [--size 640]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py --batch 128

Prints one JSON line (images/s over all ranks, ms/step, achieved conv TFLOP/s = 3 x forward FLOPs / time).
This is a secondary workload: bench.py (the driver's contract) measures the inference + NMS headline metric."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from ayolov2_b200 import dist_utils as du  # noqa: E402
from ayolov2_b200 import ops, synth  # noqa: E402
from ayolov2_b200.loss import ComputeLoss  # noqa: E402

HYP = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, label_smoothing=0.0)


def synth_targets(bs: int, seed: int) -> torch.Tensor:
    """n ~ Poisson(7) boxes per image, cls randint(80), xy U(0.05, 0.95), wh LogU(0.02, 0.6)."""
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(bs):
        n = int(rng.poisson(7))
        for _ in range(n):
            w, h = np.exp(rng.uniform(np.log(0.02), np.log(0.6), 2))
            x, y = rng.uniform(0.05, 0.95, 2)
            w, h = min(w, 2 * min(x, 1 - x)), min(h, 2 * min(y, 1 - y))
            rows.append([b, rng.integers(0, 80), x, y, w, h])
    return torch.tensor(rows, dtype=torch.float32) if rows else torch.zeros((0, 6))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128, help="GLOBAL batch (split over ranks like the reference)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--model", default="yolov5s")
    ap.add_argument("--profile", action="store_true", help="print a per-kernel table (torch profiler) of 2 extra steps")
    args = ap.parse_args()
    rank, local_rank, world = du.env_ranks()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    du.init(device=dev)
    bs = du.shard_batch(args.batch, rank, world)
    model = synth.build_model(args.model, seed=0).to(dev).train()
    model.hyp = dict(HYP)
    loss_fn = ComputeLoss(model)
    params = [p for p in model.parameters()]
    imgs = [torch.randint(0, 256, (bs, 3, args.size, args.size), dtype=torch.uint8, device=dev) for _ in range(2)]
    tgts = [synth_targets(bs, 10 * rank + i).to(dev) for i in range(2)]
    lr, momentum, ema_decay = 0.01, 0.937, 0.9999
    wd = 5e-4 * args.batch / 64
    # Flat optimizer state in the engine's gradient layout (model.parameters() order, 16-byte aligned slices): the
    # parameters become views of one fp32 buffer, so SGD-nesterov + EMA is ONE fused launch and the DDP exchange ONE
    # all-reduce. The reference's three parameter groups (yolo_trainer.py:149-168: BN weights / biases without decay,
    # other weights with decay) become a 0/1 mask on the decay term.
    model(imgs[0].float() / 255.0)  # builds the train engine (fixes the flat layout)
    eng = next(iter(model.__dict__["_train_engine_cache"].values()))
    flat_p = torch.zeros_like(eng.pg_flat)
    decay_mask = torch.zeros_like(eng.pg_flat)
    for p, o in zip(params, eng.pg_offsets):
        flat_p[o:o + p.numel()].copy_(p.data.reshape(-1))
        p.data = flat_p[o:o + p.numel()].view(p.shape)
        if p.dim() > 1:
            decay_mask[o:o + p.numel()] = 1.0
    flat_m = torch.zeros_like(flat_p)
    flat_e = flat_p.clone() if rank == 0 else None

    def step(i: int) -> float:
        x = imgs[i % 2].float() / 255.0  # prepare_img (abstract_trainer.py:252-261)
        preds = model(x)
        loss, items = loss_fn(preds, tgts[i % 2])
        if world > 1:
            loss = loss * world  # yolo_trainer.py:325-326
        loss.backward()
        g = eng.last_grad_flat  # d(loss)/d(parameters), flat; the per-parameter .grad tensors are views of it
        if world > 1:
            torch.distributed.all_reduce(g)
            g.div_(world)
        g.addcmul_(decay_mask, flat_p, value=wd)  # weight decay of the decayed group (torch.optim.SGD adds wd * p to the gradient)
        ops.sgd_ema_step(flat_p, g, flat_m, flat_e, lr, momentum, 0.0, True, ema_decay)
        for p in params:
            p.grad = None
        return items

    for i in range(args.warmup):
        items = step(i)
    du.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        items = step(i)
    e1.record()
    du.barrier()
    ms = du.max_over_ranks(e0.elapsed_time(e1), dev)
    eng = next(iter(model.__dict__["_train_engine_cache"].values()))
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for i in range(2):
                step(i)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70), file=sys.stderr)
    if rank == 0:
        ips = world * bs * args.steps / (ms / 1000.0)
        line = {"metric": "images/sec train step (fwd + ComputeLoss + bwd + SGD/EMA)", "value": ips, "unit": "images/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "dtype": "bf16 (fp32 accumulate, fp32 master weights)", "data": "synthetic",
                "config": {"workload": f"{args.model} {args.size}x{args.size} global batch {args.batch} ({bs}/GPU), "
                                       "Poisson(7) targets/img"},
                "conv_tflops_3x_fwd": 3.0 * eng.flops_fwd * world * args.steps / (ms / 1000.0) / 1e12,
                "loss_items_last": [float(v) for v in items]}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
