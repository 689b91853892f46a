"""Runs only the stride-8 detect convolution (with the fused candidate epilogue) of the benchmark model a few times, for
`ncu --set full --import-source on -k regex:conv_tc -s 2 -c 1` (source-level stall attribution of the epilogue)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from ayolov2_b200 import synth  # noqa: E402
from ayolov2_b200.detector import Detector  # noqa: E402

dev = torch.device("cuda:0")
model = bench.calibrated_model(torch.device("cuda:0"))
det = Detector(model, bench.BATCH, bench.H, bench.W, conf_thres=bench.CONF, iou_thres=bench.IOU, in_dtype=torch.uint8, device=dev)
imgs = bench.synth_images(bench.BATCH, 1000).to(dev)
det.run_device(imgs)
torch.cuda.synchronize()
level = int(sys.argv[1]) if len(sys.argv) > 1 else 0
pl = det.engine.head_plans[level]
torch.cuda.profiler.start()
for _ in range(4):
    det.nms_ws.begin_candidates()
    pl.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
