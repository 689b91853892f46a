"""Times the GPU input side (ay2_letterbox_collate) on a BASELINE-size batch (64 x 640 x 640) with CUDA events and reports
it against the measured HBM copy bandwidth. Usage: python tools/bench_input.py [mix]   (mix: copy | resize | mixed)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ayolov2_b200 import data_loader as dl, ops  # noqa: E402


def measure(mix: str, B: int = 64, H: int = 640, W: int = 640, iters: int = 50):
    rng = np.random.default_rng(0)
    shapes = []
    for i in range(B):
        kind = {"copy": 0, "resize": 1}.get(mix, i % 2)
        shapes.append((640, int(rng.integers(400, 641)) // 4 * 4) if kind == 0 else (int(rng.integers(240, 500)), int(rng.integers(240, 600))))
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    pbs = [dl.pack_batch(imgs[s:] + imgs[:s], (H, W), pin=True) for s in (0, 1)]  # two batches alternate
    dev = [torch.empty(pb.arena.numel(), dtype=torch.uint8, device="cuda") for pb in pbs]
    import dataclasses
    dpb = [dataclasses.replace(pb, arena=d.copy_(pb.arena)) for pb, d in zip(pbs, dev)]
    out = torch.empty((B, 3, H, W), dtype=torch.uint8, device="cuda")
    s2d = ops.ActView(torch.zeros((B, H // 2, W // 2 + 8, 16), dtype=torch.bfloat16, device="cuda"), 0, 16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    outs = [out, torch.empty_like(out)]
    s2ds = [s2d, ops.ActView(torch.zeros_like(s2d.buf), 0, 16)]
    reps = 10
    for name, fn in (("nchw_u8", lambda i: dpb[i & 1].to_device(out=outs[i & 1])),
                     ("s2d_bf16", lambda i: dpb[i & 1].to_space_to_depth(s2ds[i & 1], 1 / 255.0, x_offset=1))):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()  # `reps` launches back to back, two input / output sets alternating (each launch's
        with torch.cuda.graph(g):   # working set of 140-270 MB exceeds the 126 MB L2): no host time inside the timed region
            for i in range(reps):
                fn(i)
        ts = []
        for it in range(iters // 5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
        ms = float(np.median(ts))
        src = sum(3 * h * w for h, w in shapes)
        wr = B * H * W * (3 if name == "nchw_u8" else 8)  # s2d: 32 B per 2x2 pixels = 8 B / pixel
        res[name] = dict(ms=round(ms, 5), src_bytes=src, out_bytes=wr, gbs=round((src + wr) / ms / 1e6, 1))
    return res


if __name__ == "__main__":
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    for mix in (sys.argv[1:] or ["copy", "resize", "mixed"]):
        print(mix, json.dumps(measure(mix)))
    print("peaks", json.dumps(peaks)[:400])
