"""Runs N eager (non-graph) steps of the bs64 yolov5s detector for ncu: every kernel is a plain launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from ayolov2_b200 import synth  # noqa: E402
from ayolov2_b200.detector import Detector  # noqa: E402
from ayolov2_b200.nms import nms_device  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(0)
model = bench.calibrated_model(torch.device("cuda:0"))
img = torch.randint(0, 256, (batch, 3, 640, 640), dtype=torch.uint8, device="cuda")
det = Detector(model, batch, 640, 640, in_dtype=torch.uint8)
eng = det.engine
eng._img = img
torch.cuda.synchronize()
print("PROFILE_BEGIN", flush=True)
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()  # ncu --profile-from-start off: exactly ONE step is profiled
    eng.b.s2d_step()
    det._body()  # exactly the launch sequence the benchmarked CUDA graph replays (fused head -> NMS)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("PROFILE_END launches/step", det.launches_per_step(), "candidates", int(det.nms_ws.ws[:4 * batch].view(torch.int32).sum()),
      "dets", int(det.nms_ws.count.sum()))
