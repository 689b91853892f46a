"""Per-kernel NMS timing (torch profiler / CUPTI) and candidate statistics on the benchmark workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ayolov2_b200 import synth
from ayolov2_b200.detector import Detector

dev = torch.device("cuda:0")
model = synth.build_model("yolov5s", seed=0).to(dev)
with torch.no_grad():
    _, raw = model(bench.synth_images(4, 7).to(dev).float() / 255.0)
synth.calibrate_head(model, raw)
model.invalidate_engine()
det = Detector(model, bench.BATCH, bench.H, bench.W, conf_thres=bench.CONF, iou_thres=bench.IOU, in_dtype=torch.uint8, device=dev)
imgs = bench.synth_images(bench.BATCH, 1000).to(dev)
for _ in range(3):
    det.run_device(imgs)
torch.cuda.synchronize()
out = det.nms_ws.out.view(bench.BATCH, -1, 6)
cnt = det.nms_ws.count
print("dets/img min/mean/max", cnt.min().item(), cnt.float().mean().item(), cnt.max().item())
for b in range(3):
    cls = out[b, :cnt[b], 5].long()
    h = torch.bincount(cls, minlength=80)
    print(f"img{b}: kept-class histogram top5", h.topk(5))
eng = det.engine
run = lambda: det.nms_ws.run_logits(det.levels, eng.head_logits, det.conf_thres, det.iou_thres, agnostic=det.agnostic)
if os.environ.get("AY2_NMS_TRACE_ONLY"):
    for _ in range(3):
        run(); torch.cuda.synchronize()
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity
run(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10):
        run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        det.run_device(imgs)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
if os.environ.get("AY2_NMS_TRACE"):
    for _ in range(3):
        run(); torch.cuda.synchronize()
