"""Per-kernel breakdown of the training step (yolov5s 640x640, batch 128 on one GPU by default) with torch.profiler:
which kernels the step's milliseconds go to. GPU box only (tools/gpu.sh py:tools/prof_train.py).
Usage: python tools/prof_train.py [batch] [rows]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ayolov2_b200 import synth  # noqa: E402
from ayolov2_b200.trainer import TrainStep  # noqa: E402


def main() -> None:
    bs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    dev = torch.device("cuda:0")
    model = synth.build_model("yolov5s", seed=0)
    ts = TrainStep(model, bench.HYP, batch_size=bs, batches_per_epoch=1000, epochs=300, img_size=640, device=dev)
    imgs = [torch.randint(0, 256, (bs, 3, 640, 640), dtype=torch.uint8, device=dev) for _ in range(2)]
    tgts = [bench.synth_targets(bs, i).to(dev) for i in range(2)]
    for i in range(5):
        ts.training_step((imgs[i % 2], tgts[i % 2], None, None), i, 0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(10):
        ts.training_step((imgs[i % 2], tgts[i % 2], None, None), 100 + i, 0)
    b.record()
    torch.cuda.synchronize()
    print(f"train step bs{bs}: {a.elapsed_time(b) / 10:.3f} ms/step (CUDA events, 10 steps)")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for i in range(2):
            ts.training_step((imgs[i % 2], tgts[i % 2], None, None), 200 + i, 0)
        torch.cuda.synchronize()
    print("two profiled steps:")
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=rows, max_name_column_width=70))


if __name__ == "__main__":
    main()
