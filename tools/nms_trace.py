"""Steady-state phase clocks of the NMS sort/suppress kernel on the benchmark's candidates (AY2_NMS_TRACE=1 makes the
library print clock64 stamps of image 0's first CTA after every eager launch; the first launch is cold)."""
import os
import sys

os.environ["AY2_NMS_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from ayolov2_b200.detector import Detector  # noqa: E402

dev = torch.device("cuda:0")
model = bench.calibrated_model(dev)
det = Detector(model, bench.BATCH, bench.H, bench.W, conf_thres=bench.CONF, iou_thres=bench.IOU, in_dtype=torch.uint8, device=dev)
imgs = bench.synth_images(bench.BATCH, 1000).to(dev)
for _ in range(3):
    det.run_device(imgs)
torch.cuda.synchronize()
print("--- eager relaunches of the suppression kernel on the same candidate slots ---", file=sys.stderr, flush=True)
for _ in range(4):
    det.nms_ws.run_candidates(det.levels, det.engine.head_logits, det.iou_thres, agnostic=False)
    torch.cuda.synchronize()
