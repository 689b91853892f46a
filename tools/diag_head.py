"""Diagnostics of the fused head-candidate epilogue on the benchmark workload: per level, the share of rows passing the
objectness test, the per-tile (128 pixels x 3 anchors) distribution of passing pairs, and the detect convolutions' time
with and without the candidate epilogue armed."""
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
for _ in range(3):
    det.run_device(imgs)
torch.cuda.synchronize()
eng = det.engine
for lv, lg in enumerate(eng.head_logits):
    t = lg.tensor().float()[..., :255].reshape(bench.BATCH, lg.H * lg.W, 3, 85)
    obj = torch.sigmoid(t[..., 4])
    conf = obj * torch.sigmoid(t[..., 5:]).max(-1).values
    passed = obj > bench.CONF
    cand = passed & (conf > bench.CONF)
    flat = passed.reshape(-1, 3)
    ntile = flat.shape[0] // 128
    per_tile = flat[:ntile * 128].reshape(ntile, 128 * 3).sum(1).float()
    print(f"level {lv}: rows passing obj {float(passed.float().mean()):.4f}, candidates {float(cand.float().mean()):.4f} "
          f"({int(cand.sum()) // bench.BATCH}/img); per-tile passing pairs: mean {float(per_tile.mean()):.1f}, "
          f"share of tiles with > 256: {float((per_tile > 256).float().mean()):.3f}, with 0: {float((per_tile == 0).float().mean()):.3f}, "
          f"max {int(per_tile.max())}")


def time_plan(pl, n=20):
    det.nms_ws.begin_candidates()
    pl.run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        det.nms_ws.begin_candidates()
        pl.run()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for lv, (pl, off) in enumerate(zip(eng.head_plans, eng.head_row_off)):
    t_on = time_plan(pl)
    pl.set_head_candidates(None)
    t_off = time_plan(pl)
    pl.set_head_candidates(det.nms_ws, eng.na, off)
    print(f"level {lv}: detect conv {t_on:.1f} us with the candidate epilogue, {t_off:.1f} us without; pair={pl.pair}")
