#!/usr/bin/env python
"""bench.py — images/sec of the YOLOv5s forward + NMS hot path (BASELINE.json `metric`, configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of 64 synthetic 640x640 RGB images per GPU:
uint8 -> /255 -> yolov5s forward (bf16 tensor cores, fp32 accumulate) -> decode -> batched NMS (conf 0.25, iou 0.45).
  value : whole-job images/s with the input batch already resident in HBM (CUDA-event timed, max over ranks)
  e2e   : the same metric through the public host API (Detector.submit/collect): pinned-host uint8 batch ->
          H2D -> kernels -> D2H of the detections, every step, copies inside the timed region
  roofline : the conv kernel family: a CUDA graph holding exactly the step's convolution launches, replayed and timed with
             CUDA events (the in-graph duration, <= ms_per_step), against the measured bf16 peak
  extras   : the other BASELINE.json configs on the same box -- train_step (configs[2]: yolov5s bs128 global fwd + ComputeLoss
             + bwd + gradient all-reduce + SGD/EMA, with the all-reduce's exposed time), tucker (configs[3]: decomposed vs
             dense images/s and the logits error of the fused chain), yolov5l_train (configs[4], 8 GPUs), nms_synthetic
             (SURVEY 8(d) tensor)
  cpu_baseline : the CPU oracle (fp32 PyTorch restatement + NMS restatement) on this box's host cores
`--impl reference` times that same CPU path as the reference arm (the reference's operators live in the
un-vendored `kindle` package, so the restatement in oracle/ is the closest runnable form).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec @640 bs64 yolov5s fwd+NMS"
GFLOP_PER_IMG = 16.4336  # yolov5s @640 conv FLOPs (2*MAC) per image
ACT_MB_PER_IMG = 121.6   # bf16 activation traffic per image, layer by layer
BATCH, H, W = 64, 640, 640
CONF, IOU = 0.25, 0.45


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8:
                self.rows.append(f)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def synth_images(batch: int, seed: int):
    import torch

    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, 3, H, W), generator=g, dtype=torch.uint8)


_THREADS = None


def pick_threads() -> int:
    """Host threads for the CPU path: the count (<= visible cores) that runs a small forward fastest. A box may
    expose more logical CPUs than its cgroup lets us use; oversubscribing OpenMP there is catastrophically slow, and
    the reference arm is supposed to use the host as well as it can."""
    global _THREADS
    if _THREADS is not None:
        return _THREADS
    import torch

    from ayolov2_b200 import synth as model_utils
    from oracle import yolo_oracle

    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count() or 1
    model = model_utils.build_model("yolov5s", seed=0)
    x = torch.rand(2, 3, 320, 320)
    best, best_t = 1, 1e30
    cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu} | {min(ncpu, 8)})
    for c in cands:
        torch.set_num_threads(c)
        yolo_oracle.forward(model, x)
        t0 = time.perf_counter()
        yolo_oracle.forward(model, x)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
        if dt > 4 * best_t:
            break
    _THREADS = best
    return best


def cpu_path_images_per_s(n_batches: int, bs: int, threads: int, min_seconds: float = 0.0):
    """The reference's CPU path for this metric: fp32 forward (oracle restatement of the kindle operators) +
    non_max_suppression restatement, `threads` host threads, uint8 -> /255 included."""
    import torch

    from ayolov2_b200 import synth as model_utils
    from oracle import nms_oracle, yolo_oracle

    torch.set_num_threads(threads)
    model = model_utils.build_model("yolov5s", seed=0)
    imgs = synth_images(bs, 123)
    x = imgs.float() / 255.0
    yolo_oracle.forward(model, x[:1])  # warm-up
    t0 = time.perf_counter()
    done = 0
    while done < n_batches or time.perf_counter() - t0 < min_seconds:  # at least n_batches, and at least min_seconds of CPU work
        x = imgs.float() / 255.0
        pred, _ = yolo_oracle.forward(model, x)
        nms_oracle.non_max_suppression(pred, CONF, IOU)
        done += 1
    dt = time.perf_counter() - t0
    return done * bs / dt, dt


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    threads = pick_threads()
    bs = 8
    per_step = 1  # one bounded sample (8 images) per step
    # warm-up
    for _ in range(max(args.warmup, 1)):
        cpu_path_images_per_s(1, bs, threads)
    t0 = time.perf_counter()
    ips_list = []
    for _ in range(args.steps):
        ips, _dt = cpu_path_images_per_s(per_step, bs, threads)
        ips_list.append(ips)
    total = time.perf_counter() - t0
    value = statistics.median(ips_list)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"yolov5s 640x640 fwd+NMS conf {CONF} iou {IOU}; CPU path, bounded sample of {bs} images per step"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{bs} images per step x {args.steps} steps, fp32 PyTorch CPU restatement of the kindle operators "
                                   "+ reference-pinned NMS restatement (kindle itself is not installable: no reference wheel)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process to the CPUs closest to its GPU (NVML affinity) BEFORE any pinned host memory is allocated, so the
    staging buffers of the end-to-end leg are first-touched on the GPU's NUMA node. Returns a short description."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(cpus & allowed)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"cpus {cpus[0]}-{cpus[-1]} ({len(cpus)}) by NVML affinity"
        return "NVML affinity empty within the cgroup: not bound"
    except Exception as e:  # no NVML / not permitted: run unbound
        return f"not bound ({type(e).__name__})"


def time_graph(body, steps: int, warmup: int = 3) -> float:
    """ms per replay of a CUDA graph capturing `body` (CUDA events on the replay stream)."""
    import torch

    body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    for _ in range(warmup):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


HYP = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, lrf=0.1, momentum=0.937,
           weight_decay=5e-4, warmup_epochs=3.0, warmup_momentum=0.8, warmup_bias_lr=0.1,
           optimizer_params=dict(lr=0.01, momentum=0.937, nesterov=True))  # res/configs/cfg/train_config.yaml:29-54


def synth_targets(bs: int, seed: int):
    """n ~ Poisson(7) boxes per image, cls randint(80), xy U(0.05, 0.95), wh LogU(0.02, 0.6) clipped."""
    import numpy as np
    import torch

    rng = np.random.default_rng(seed)
    rows = []
    for b in range(bs):
        for _ in range(int(rng.poisson(7))):
            w, h = np.exp(rng.uniform(np.log(0.02), np.log(0.6), 2))
            x, y = rng.uniform(0.05, 0.95, 2)
            rows.append([b, rng.integers(0, 80), x, y, min(w, 2 * min(x, 1 - x)), min(h, 2 * min(y, 1 - y))])
    return torch.tensor(rows, dtype=torch.float32) if rows else torch.zeros((0, 6))


def calibrated_model(dev, name: str = "yolov5s"):
    """The benchmark's model: random-init `name` (seed 0) whose detect head is re-scaled so that the NMS leg is not vacuous:
    ~8 % of the rows of every pyramid level (~2,000 of 25,200 per image) are candidates at conf 0.25, over
    many classes (ayolov2_b200.synth.calibrate_head). The profiling tools under tools/ use the same function."""
    from ayolov2_b200 import synth as model_utils

    model = model_utils.build_model(name, seed=0).to(dev)
    sample = synth_images(32, 7).to(dev).float() / 255.0
    model_utils.calibrate_head(model, lambda: model(sample)[1], cand_frac=0.08)
    return model


def bench_train_step(name: str, global_batch: int, size: int, steps: int, warmup: int, rank: int, world: int, dev, barrier,
                     max_over_ranks):
    """One BASELINE train config through ayolov2_b200.trainer.TrainStep (the drop-in of YoloTrainer.training_step): uint8
    batch on the device -> /255 -> forward (batch-statistics BN) -> ComputeLoss -> backward -> bucketed gradient all-reduce
    overlapped with the backward -> fused SGD-nesterov + EMA. Returns the extras entry (rank 0) or None."""
    import torch

    from ayolov2_b200 import synth as model_utils
    from ayolov2_b200.trainer import TrainStep

    bs = global_batch // world
    model = model_utils.build_model(name, seed=0)
    ts = TrainStep(model, HYP, batch_size=global_batch, batches_per_epoch=1000, epochs=300, img_size=size, device=dev)
    imgs = [torch.randint(0, 256, (bs, 3, size, size), dtype=torch.uint8, device=dev) for _ in range(2)]
    tgts = [synth_targets(bs, 10 * rank + i).to(dev) for i in range(2)]

    def run(n: int, first: int) -> float:
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            ts.training_step((imgs[i % 2], tgts[i % 2], None, None), first + i, 0)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / n

    run(max(warmup, 4), 0)  # eager pass, graph capture, first replays
    ms = run(steps, 100)
    exposed = None
    if world > 1:  # the same steps without the exchange: what the all-reduce adds to the step (its EXPOSED time)
        ts.skip_exchange = True
        ms_local = run(steps, 200)
        ts.skip_exchange = False
        exposed = ms - ms_local
    if rank != 0:
        return None
    eng = ts._engine()
    return {"config": f"{name} {size}x{size} global batch {global_batch} ({bs}/GPU), Poisson(7) targets/img, SGD-nesterov + EMA",
            "images_per_s": world * bs / (ms / 1000.0), "ms_per_step": ms, "n_gpus": world,
            "conv_tflops_per_gpu_3x_fwd": 3.0 * eng.flops_fwd / (ms / 1000.0) / 1e12,
            "allreduce": None if world == 1 else {"bytes": int(eng.pg_flat.numel() * 4), "buckets": len(getattr(eng, "bwd_chunks", [0])),
                                                  "exposed_ms": exposed, "overlapped_with_backward": len(getattr(eng, "bwd_chunks", [])) > 1},
            "loss_items_last": [float(v) for v in ts.mloss.tolist()]}


def bench_tucker(steps: int, dev):
    """BASELINE configs[3]: Tucker-2 decomposed yolov5s (fixed ranks ceil(0.5 C)) bs64 640x640 vs dense, one GPU."""
    import torch

    import ayolov2_b200
    from ayolov2_b200 import engine as eng_mod, synth as model_utils, tucker

    img = torch.randint(0, 256, (BATCH, 3, H, W), dtype=torch.uint8, device=dev)
    out = {}

    def run(model, label, fuse=True):
        eng_mod.Builder.FUSE_CHAINS = fuse
        try:
            e = eng_mod.Engine(model, BATCH, H, W, in_dtype=torch.uint8, scale=1 / 255.0, want_raw=False, use_graph=True)
        finally:
            eng_mod.Builder.FUSE_CHAINS = True
        for _ in range(3):
            e.run(img)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            e.run(img)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        out[label] = {"ms_per_step": ms, "images_per_s": BATCH / ms * 1000.0, "launches": len(e.b.steps),
                      "fused_chain_launches": sum(1 for p in e.b.plans if type(p).__name__ == "ChainPlan"),
                      "params": sum(p.numel() for p in model.parameters())}
        return e

    dense = model_utils.build_model("yolov5s", seed=0).to(dev).eval()
    run(dense, "dense")
    dec = model_utils.build_model("yolov5s", seed=0)
    n = len(tucker.decompose_model_fixed(dec, ratio=0.5))
    dec = dec.to(dev).eval()
    run(dec, "tucker_fused_chain")
    run(dec, "tucker_three_launches", fuse=False)
    # logits of the fused bf16 path against the split-precision (fp32-equivalent) evaluation of the same chains
    x = torch.rand((2, 3, 320, 320), device=dev)
    with torch.no_grad():
        _, raw_b = dec(x)
        ayolov2_b200.set_precision(dec, "bf16x3")
        _, raw_p = dec(x)
        ayolov2_b200.set_precision(dec, "bf16")
    err = max(float((a - b).norm() / b.norm()) for a, b in zip(raw_b, raw_p))
    out.update({"decomposed_convs": n, "rank_ratio": 0.5, "fused_vs_dense": out["tucker_fused_chain"]["images_per_s"] / out["dense"]["images_per_s"],
                "bf16_logits_rel_l2_vs_fp32_equivalent": err,
                "note": "fp32-equivalent (bf16x3) evaluation vs the nn.Sequential oracle: 3e-6 rel-L2 (tests/test_precise_gpu.py)"})
    return out


def bench_nms_synthetic(dev, steps: int = 20):
    """NMS alone on the synthetic tensor SURVEY.md 8(d) prescribes (64 x 25200 x 85, ~2,000 clustered candidates / image)."""
    import torch

    from ayolov2_b200 import synth as model_utils
    from ayolov2_b200.nms import nms_device

    pred = model_utils.synth_predictions(BATCH, 25200, 80, seed=0, device=dev)
    ws = nms_device(pred, CONF, IOU)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        nms_device(pred, CONF, IOU, workspace=ws)
    b.record()
    torch.cuda.synchronize()
    return {"ms_dense_3_launches": a.elapsed_time(b) / steps, "candidates": int(ws.ws[:4 * BATCH].view(torch.int32).sum().item()),
            "detections": int(ws.count.sum().item()),
            "config": "(64, 25200, 85) fp32, Bernoulli(0.08) x U(0.25, 1) objectness, 200 box clusters / image, conf 0.25 iou 0.45"}


def bench_input_side(det, dev, rank: int, world: int, barrier, max_over_ranks, steps: int = 30):
    """the loader's letterbox + channel flip + collate (scripts/data_loader/data_loader.py:380-393,
    461-477) on the GPU. (1) the kernel alone, fused into the stem's space-to-depth input, against the HBM roofline;
    (2) end to end through Detector.submit_packed from LOADED images in pinned host memory (ragged HWC BGR uint8, long side
    640 like `_load_image` leaves them), next to the CPU oracle's letterbox + collate of the same images."""
    import dataclasses

    import numpy as np
    import torch

    from ayolov2_b200 import data_loader as dl
    from ayolov2_b200 import ops
    from oracle import input_oracle

    rng = np.random.default_rng(0)
    out = {}
    # validation-like mix: long side 640 (3:4, 4:3, 2:3 aspect ratios and squares), a quarter smaller images that are up-scaled
    def shapes_of(kind):
        res = []
        for i in range(BATCH):
            if kind == "copy" or (kind == "val" and i % 4):
                res.append([(640, 480), (480, 640), (640, 428), (428, 640), (640, 640)][i % 5])
            else:
                res.append((int(rng.integers(240, 500)), int(rng.integers(240, 600))))
        return res
    peak = _peaks().get("hbm_gbs") or 6532.2
    s2d = [ops.ActView(torch.zeros_like(det.engine.b.s2d_view.buf), 0, 16) for _ in range(2)]
    for kind in ("copy", "resize", "val"):
        shp = shapes_of(kind)
        imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shp]
        pbs = [dl.pack_batch(imgs[k:] + imgs[:k], (H, W), pin=True) for k in (0, 1)]
        dpb = [dataclasses.replace(pb, arena=pb.arena.to(dev)) for pb in pbs]
        reps = 10

        def body():
            for i in range(reps):  # two input / output sets alternate; one launch moves > 126 MB (L2)
                dpb[i & 1].to_space_to_depth(s2d[i & 1], 1.0 / 255.0, x_offset=1)
        ms = time_graph(body, steps) / reps
        src = sum(3 * h * w for h, w in shp)
        wr = BATCH * H * W * 8
        out[f"kernel_{kind}"] = {"ms": ms, "algorithmic_bytes": src + wr, "achieved_gbs": (src + wr) / ms / 1e6,
                                 "frac_of_hbm_peak": (src + wr) / ms / 1e6 / peak}
        if kind != "val":
            continue
        # end to end from loaded images. The two collectives below run exactly once on every rank whatever happens locally
        # (a rank that failed reports an infinite time instead of leaving the others in a barrier)
        err, dt_local = None, float("inf")
        try:
            for i in range(3):
                det.collect(det.submit_packed(pbs[i & 1]))
            torch.cuda.synchronize()
        except Exception as e:
            err = e
        barrier()  # every rank stages its own batches at the same time (the host -> device link is the shared resource)
        if err is None:
            try:
                t0 = time.perf_counter()
                pending = []
                for i in range(steps):
                    pending.append(det.submit_packed(pbs[i & 1]))
                    if len(pending) > det.slots - 1:
                        det.collect(pending.pop(0))
                while pending:
                    det.collect(pending.pop(0))
                dt_local = time.perf_counter() - t0
            except Exception as e:
                err = e
        dt = max_over_ranks(dt_local * 1000.0) / 1000.0
        if err is not None or dt == float("inf"):
            out["e2e_from_loaded_images"] = {"error": f"{type(err).__name__}: {err}"[:200] if err else "another rank failed"}
        else:
            out["e2e_from_loaded_images"] = {"images_per_s": world * BATCH * steps / dt, "ms_per_step": dt * 1000 / steps, "n_gpus": world,
                                             "h2d_bytes_per_step": int(pbs[0].arena.numel()),
                                             "padded_batch_bytes": BATCH * 3 * H * W,
                                             "h2d_gbs_per_rank": int(pbs[0].arena.numel()) / (dt / steps) / 1e9}
        if rank != 0:
            continue
        t0 = time.perf_counter()
        n = 16
        input_oracle.load_and_collate(imgs[:n], (H, W))
        dt = time.perf_counter() - t0
        out["cpu_oracle"] = {"images_per_s": n / dt, "cores": 1, "kind": "port", "sample": f"{n} images of the same mix, numpy restatement of _letterbox + collate_fn"}
        try:  # the reference's own dependency for this step, when the box has it: cv2.resize + copyMakeBorder + transpose + stack
            import cv2

            def cv2_letterbox(im):
                (uw, uh), _, _, (top, bottom, left, right) = dl.letterbox_geometry(im.shape[:2], (H, W), auto=False)
                if (uh, uw) != im.shape[:2]:
                    im = cv2.resize(im, (uw, uh), interpolation=cv2.INTER_LINEAR)
                im = cv2.copyMakeBorder(im, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))
                return np.ascontiguousarray(im.transpose((2, 0, 1))[::-1])
            cv2.setNumThreads(1)
            t0 = time.perf_counter()
            reps = 4
            for _ in range(reps):
                np.stack([cv2_letterbox(im) for im in imgs], 0)
            dt = time.perf_counter() - t0
            out["cpu_cv2"] = {"images_per_s": reps * len(imgs) / dt, "cores": 1, "kind": "reference dependency (opencv " + cv2.__version__ + ")",
                              "sample": f"{reps} x {len(imgs)} images of the same mix: cv2.resize + copyMakeBorder + transpose + np.stack"}
        except Exception as e:  # no cv2 on the box (or anything else): the entry is optional
            out["cpu_cv2"] = {"unavailable": f"{type(e).__name__}: {e}"[:120]}
    out["config"] = f"{BATCH} loaded images -> ({BATCH}, 3, {H}, {W}); 'val': long side 640 mixed aspect ratios, one in four up-scaled; fused into the stem's space-to-depth input (bf16, /255)"
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the train / Tucker / synthetic-NMS extras")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from ayolov2_b200 import _lib
    from ayolov2_b200.detector import Detector
    from ayolov2_b200 import synth as model_utils

    args.warmup = max(args.warmup, 3)  # timing rules: at least 3 warm-up steps (the line reports the count actually run)
    args.steps = max(args.steps, 1)
    numa = bind_to_gpu_numa_node(local_rank)  # before the pinned staging buffers exist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    model = calibrated_model(dev)
    det = Detector(model, BATCH, H, W, conf_thres=CONF, iou_thres=IOU, in_dtype=torch.uint8, device=dev)
    host_imgs = [synth_images(BATCH, 1000 + rank * 10 + i).pin_memory() for i in range(3)]
    dev_imgs = [h.to(dev) for h in host_imgs[:2]]

    # ------------------------------------------------------------------ device-resident throughput (`value`)
    for i in range(args.warmup):
        det.run_device(dev_imgs[i % 2])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_launch0 = _lib.launch_count()
    e0.record()
    for i in range(args.steps):
        det.run_device(dev_imgs[i % 2])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = world * BATCH * args.steps / (ms / 1000.0)
    ndet = int(det.nms_ws.count.sum().item())
    ncand = int(det.nms_ws.ws[:4 * BATCH].view(torch.int32).sum().item())  # per-image candidate counters

    # ------------------------------------------------------------------ end to end through the host API (`e2e`)
    for i in range(args.warmup):
        det.collect(det.submit(host_imgs[i % 3]))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    pending = []
    for i in range(args.steps):
        pending.append(det.submit(host_imgs[i % 3]))
        if len(pending) > det.slots - 1:  # slots - 1 batches in flight: H2D of the next and D2H of the previous overlap the kernels
            det.collect(pending.pop(0))
    while pending:
        det.collect(pending.pop(0))
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0))
    e2e_value = world * BATCH * args.steps / (e2e_ms / 1000.0)
    h2d = host_imgs[0].numel() * host_imgs[0].element_size()
    d2h = det.host_out[0].numel() * 4 + det.host_cnt[0].numel() * 4
    n_slots, launches_per_step = det.slots, det.launches_per_step()

    # ------------------------------------------------------------------ roofline of the conv kernel family
    peaks = _peaks()
    roof = None
    if rank == 0:
        eng = det.engine
        evs = []
        for rep in range(3):  # eager replays of a full step, every conv launch bracketed by events
            eng._img = dev_imgs[rep % 2]
            if det.fused_candidates:
                det.nms_ws.begin_candidates()  # the detect convolutions below append this step's NMS candidates
            for s in eng.b.steps:
                is_conv = getattr(s, "__self__", None) is not None and s.__self__.__class__.__name__ in ("ConvPlan", "ChainPlan")
                if is_conv:
                    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    s()
                    b_.record()
                    if rep > 0:
                        evs.append((s.__self__, a, b_))
                else:
                    s()
        torch.cuda.synchronize()
        per_plan = {}
        for plan, a, b_ in evs:
            per_plan.setdefault(id(plan), [plan, 0.0, 0])
            per_plan[id(plan)][1] += a.elapsed_time(b_)
            per_plan[id(plan)][2] += 1
        conv_ms_eager = sum(v[1] / v[2] for v in per_plan.values())  # per-launch events: includes the launch gaps a graph hides
        # The in-step duration of the conv family: a CUDA graph holding EXACTLY the step's convolution launches (the same
        # plans, in order, plus the memset that re-arms the candidate counters), replayed back to back and timed with CUDA
        # events on the replay stream. Inter-kernel gaps are the graph's own, so this is <= ms_per_step by construction.
        conv_steps = [s for s in eng.b.steps if getattr(s, "__self__", None) is not None
                      and s.__self__.__class__.__name__ in ("ConvPlan", "ChainPlan")]

        def conv_only():
            if det.fused_candidates:
                det.nms_ws.begin_candidates()
            for s in conv_steps:
                s()

        conv_ms = time_graph(conv_only, 20)

        def nms_only():
            if det.fused_candidates:
                det.nms_ws.run_candidates(det.levels, eng.head_logits, det.iou_thres, agnostic=det.agnostic)
            else:
                det.nms_ws.run_logits(det.levels, eng.head_logits, det.conf_thres, det.iou_thres, agnostic=det.agnostic)

        det.run_device(dev_imgs[0])  # leave one real step's candidates in the workspace
        torch.cuda.synchronize()
        nms_ms = time_graph(nms_only, 20)
        flops = GFLOP_PER_IMG * 1e9 * BATCH
        achieved = flops / (conv_ms / 1000.0) / 1e12
        peak = peaks["tflops_sustained"]
        abytes = ACT_MB_PER_IMG * 1e6 * BATCH
        traffic, traffic_src = None, None  # DRAM bytes of the same launches: from the committed ncu capture of this kernel state
        tfiles = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_conv_traffic.json")) \
            if os.path.isdir(os.path.join(ROOT, "profiles")) else []
        if tfiles:
            traffic = json.load(open(os.path.join(ROOT, "profiles", tfiles[-1])))["dram_bytes"]
            traffic_src = "profiles/" + tfiles[-1] + " (ncu dram__bytes_read.sum + dram__bytes_write.sum over one step's conv launches)"
        n_chain = sum(1 for v in per_plan.values() if v[0].__class__.__name__ == "ChainPlan")
        n_halo = sum(1 for v in per_plan.values() if getattr(v[0], "halo", False))
        roof = {"bound": "tensor",
                "kernel": f"conv family: {len(per_plan) - n_chain - n_halo} conv_tc_kernel + {n_halo} conv_halo_kernel + {n_chain} "
                          "conv_chain_kernel launches (60 convolutions) = one step", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_src": traffic_src,
                "peak_src": peaks["src"] + " (sustained)", "frac_of_burst_peak": achieved / peaks["tflops_burst"],
                "conv_ms_per_step": conv_ms, "conv_ms_how": "CUDA-graph replay of the step's conv launches only, CUDA events",
                "conv_ms_eager_per_launch_events": conv_ms_eager,
                "hbm_view": {"achieved_gbs": abytes / (conv_ms / 1000.0) / 1e9, "peak_gbs": peaks["hbm_gbs"],
                             "frac": abytes / (conv_ms / 1000.0) / 1e9 / peaks["hbm_gbs"]}}
        roof["nms_ms_per_step"] = nms_ms
        roof["nms_share_of_step"] = nms_ms / ms_per_step
        layers = []
        for plan, tot, n in per_plan.values():
            d = plan.desc
            if plan.__class__.__name__ == "ChainPlan":
                layers.append({"chain": [d.cin, d.c1, d.c2, d.c3], "k": d.kh, "s": d.stride, "hw": [d.in_h, d.in_w],
                               "ms": tot / n, "tflops": plan.flops / (tot / n / 1000.0) / 1e12})
                continue
            layers.append({"cin": d.cin, "cout": d.cout, "k": d.kh, "s": d.stride, "hw": [d.out_h, d.out_w],
                           "ms": tot / n, "tflops": plan.flops / (tot / n / 1000.0) / 1e12})
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(layers, open(os.path.join(ROOT, "gpurun_out", "conv_layers.json"), "w"), indent=1)

    # ------------------------------------------------------------------ the other BASELINE configs (extras)
    extras = {}
    only = [x for x in os.environ.get("AY2_BENCH_EXTRAS", "").split(",") if x]  # e.g. AY2_BENCH_EXTRAS=input_side (same on every rank)
    want = lambda name: not only or name in only  # noqa: E731
    def extra(name, fn, ranks_all=True):
        """Run one extra; an extra must never take the headline line down."""
        if args.no_extras or not want(name) or (not ranks_all and rank != 0):
            return
        try:
            extras[name] = fn()
        except Exception as e:
            extras[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    extra("train_step", lambda: bench_train_step("yolov5s", 128, 640, 20, 5, rank, world, dev, barrier, max_over_ranks))
    if world == 8 or os.environ.get("AY2_BENCH_YOLOV5L") == "1":
        extra("yolov5l_train", lambda: bench_train_step("yolov5l", 32, 640, 10, 4, rank, world, dev, barrier, max_over_ranks))
    # all ranks: the end-to-end part measures the shared host -> device path
    extra("input_side", lambda: bench_input_side(det, dev, rank, world, barrier, max_over_ranks))
    extra("tucker", lambda: bench_tucker(20, dev), ranks_all=False)

    def nms_synth():
        r = bench_nms_synthetic(dev)
        r["share_of_fwd_plus_nms"] = r["ms_dense_3_launches"] / (ms_per_step + r["ms_dense_3_launches"])
        return r
    extra("nms_synthetic", nms_synth, ranks_all=False)
    if not args.no_extras:
        barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = pick_threads()
        ips, dt = cpu_path_images_per_s(3, 8, threads, min_seconds=12.0)
        cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"{round(ips * dt)} images in batches of 8 ({dt:.1f}s): fp32 CPU oracle forward + NMS restatement; "
                         f"{threads} threads picked by a timing sweep over the {os.cpu_count()} visible CPUs"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"yolov5s.yaml 640x640 bs{BATCH}/GPU uint8 input, fwd + decode + NMS conf {CONF} iou {IOU}, "
                                   "random-init weights (seed 0), 1 process per GPU, replicas (no data-path collective)",
                       "l2": "per-step working set (79 MB input + >3 GB activations) exceeds the 126 MB L2; two input batches alternate",
                       "detections_last_step": ndet, "nms_candidates_last_step": ncand},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "h2d_gbs_per_rank": h2d / (e2e_ms / args.steps) / 1e6,
                    "pipeline": f"{n_slots} device input slots, {n_slots - 1} batches in flight; host staging: {numa}"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof,
            "extras": extras,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
