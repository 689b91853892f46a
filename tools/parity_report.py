"""Measured forward errors of the CUDA path against the fp32 CPU oracle (max-norm and rel-L2 per output), written through
tests/_parity.record. Run on the GPU box: `python tools/parity_report.py [--full]`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _parity  # noqa: E402
from ayolov2_b200 import synth  # noqa: E402
from oracle import yolo_oracle  # noqa: E402


def one(name, B, hw, seed=0, tag=""):
    model = synth.build_model(name, seed=seed)
    x = torch.rand((B, 3, *hw), generator=torch.Generator().manual_seed(5))
    want_pred, want_raw = yolo_oracle.forward(model, x)
    got_pred, got_raw = model.cuda()(x.cuda())
    torch.cuda.synchronize()
    for i, (g, w) in enumerate(zip(got_raw, want_raw)):
        _parity.record(f"report/{name}_{hw[0]}x{hw[1]}_b{B}{tag}/logits_P{i + 3}", **_parity.errs(g, w))
    gp = got_pred.float().cpu()
    _parity.record(f"report/{name}_{hw[0]}x{hw[1]}_b{B}{tag}/pred_scores", **_parity.errs(gp[..., 4:], want_pred[..., 4:]),
                   max_abs=float((gp[..., 4:] - want_pred[..., 4:]).abs().max()))
    _parity.record(f"report/{name}_{hw[0]}x{hw[1]}_b{B}{tag}/pred_boxes", **_parity.errs(gp[..., :4], want_pred[..., :4]))


if __name__ == "__main__":
    one("yolov5s", 1, (640, 640))
    one("yolov5s", 2, (320, 320))
    one("yolov5n", 2, (192, 192))
    one("yolov5_v5", 2, (256, 256))
    if "--full" in sys.argv:
        one("yolov5m", 1, (640, 640))
