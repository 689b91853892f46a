#!/usr/bin/env python
"""BASELINE.json configs[3]: Tucker-2 decomposed yolov5s (fixed ranks ceil(ratio*C)) inference bs64 640x640 on one B200,
fused chain kernel vs three launches per chain vs the dense model. Prints one JSON line per variant."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ayolov2_b200 import engine as eng_mod, synth, tucker  # noqa: E402

B, H, W = 64, 640, 640
ratio = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
steps = 30


def run(model, fuse, label):
    eng_mod.Builder.FUSE_CHAINS = fuse
    e = eng_mod.Engine(model, B, H, W, in_dtype=torch.uint8, scale=1 / 255.0, want_raw=False, use_graph=True)
    img = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, device="cuda")
    for _ in range(5):
        e.run(img)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        e.run(img)
    b_.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b_) / steps
    nchain = sum(1 for p in e.b.plans if type(p).__name__ == "ChainPlan")
    print(json.dumps({"variant": label, "ms_per_step": ms, "images_per_s": B / ms * 1000, "launches": len(e.b.steps),
                      "chain_launches": nchain, "params": sum(p.numel() for p in model.parameters())}), flush=True)
    eng_mod.Builder.FUSE_CHAINS = True
    return e


dense = synth.build_model("yolov5s", seed=0).cuda().eval()
run(dense, True, "dense")
dec = synth.build_model("yolov5s", seed=0)
names = tucker.decompose_model_fixed(dec, ratio=ratio)
dec = dec.cuda().eval()
run(dec, False, f"tucker ratio {ratio}: 3 launches per chain")
run(dec, True, f"tucker ratio {ratio}: fused chain kernel")
