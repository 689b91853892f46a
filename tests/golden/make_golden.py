"""Generates the golden fixtures in this directory by running the UNMODIFIED reference code
(/root/reference, imported through oracle/ref_import.py). Run in the build container only:

    python tests/golden/make_golden.py

  nms_golden.npz   : synthetic predictions + outputs of scripts/utils/metrics.py::non_max_suppression and
                     scripts/utils/nms.py::batched_nms for several settings
  loss_golden.npz  : seeded head outputs / targets + scripts/loss/losses.py::ComputeLoss value, items and gradients
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import nms_oracle, ref_import  # noqa: E402

NMS_SETTINGS = [
    dict(conf_thres=0.25, iou_thres=0.45),
    dict(conf_thres=0.25, iou_thres=0.45, multi_label=True),
    dict(conf_thres=0.1, iou_thres=0.65, agnostic=True),
    dict(conf_thres=0.25, iou_thres=0.45, classes=[1, 3]),
    dict(conf_thres=0.25, iou_thres=0.45, max_det=10),
    # non-default nms_type values (metrics.py:388-431); index selection exact, merged boxes / decayed scores to fp32 rounding
    dict(conf_thres=0.05, iou_thres=0.6, nms_type="batched_nms"),
    dict(conf_thres=0.25, iou_thres=0.45, nms_type="fast_nms"),
    dict(conf_thres=0.25, iou_thres=0.45, nms_type="matrix_nms"),
    dict(conf_thres=0.25, iou_thres=0.45, nms_type="merge_nms", multi_label=True),
]


def make_nms(ref):
    pred = nms_oracle.synth_predictions(3, n=400, nc=20, seed=11, cand_frac=0.3, clusters=12, img=320.0)
    pred[2, :, 4] = 0.0  # an image without candidates
    out = {"pred": pred.numpy()}
    for si, kw in enumerate(NMS_SETTINGS):
        res = ref.non_max_suppression(pred.clone(), **kw)
        for i, r in enumerate(res):
            out[f"nms{si}_img{i}"] = r.numpy()
    res = ref.batched_nms(pred.clone(), 0.05, 0.65, 100, False)
    for i, r in enumerate(res):
        out[f"bnms_img{i}"] = r.numpy()
    np.savez_compressed(os.path.join(HERE, "nms_golden.npz"), **out)
    print("nms_golden.npz", {k: v.shape for k, v in out.items() if k != "pred"})


HYP = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, label_smoothing=0.0)
ANCHORS = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]


class FakeHead:
    def __init__(self, nc):
        self.nl, self.na, self.nc = 3, 3, nc
        self.stride = torch.tensor([8.0, 16.0, 32.0])
        self.anchors = torch.tensor(ANCHORS).float().view(3, 3, 2) / self.stride.view(-1, 1, 1)


class FakeModel(torch.nn.Module):
    def __init__(self, nc, hyp):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.hyp = dict(hyp)
        self.model = [FakeHead(nc)]


def loss_inputs(seed=0, bs=3, nc=6, img=128, nt=14):
    g = torch.Generator().manual_seed(seed)
    preds = [torch.randn(bs, 3, img // s, img // s, nc + 5, generator=g) for s in (8, 16, 32)]
    t = torch.zeros(nt, 6)
    t[:, 0] = torch.randint(0, bs, (nt,), generator=g).float()
    t[:, 1] = torch.randint(0, nc, (nt,), generator=g).float()
    t[:, 2:4] = 0.05 + 0.9 * torch.rand(nt, 2, generator=g)
    t[:, 4:6] = torch.exp(np.log(0.03) + (np.log(0.7) - np.log(0.03)) * torch.rand(nt, 2, generator=g))
    return preds, t


def make_loss(ref):
    out = {}
    for case, (hyp_over, nt) in enumerate([({}, 14), ({"label_smoothing": 0.1, "cls_pw": 0.7, "obj_pw": 1.3}, 9), ({}, 0),
                                            ({"fl_gamma": 1.5, "label_smoothing": 0.05, "obj_pw": 1.2}, 11)]):
        hyp = dict(HYP, **hyp_over)
        nc = 6
        preds, targets = loss_inputs(seed=case, nc=nc, nt=nt)
        preds = [p.requires_grad_(True) for p in preds]
        model = FakeModel(nc, hyp)
        with ref_import.clamp_compat():
            loss_fn = ref.ComputeLoss(model)
            loss, items = loss_fn(preds, targets)
        loss.backward()
        out[f"c{case}_hyp"] = np.array([hyp[k] for k in sorted(hyp)], dtype=np.float64)
        out[f"c{case}_targets"] = targets.numpy()
        for i, p in enumerate(preds):
            out[f"c{case}_pred{i}"] = p.detach().numpy()
            out[f"c{case}_grad{i}"] = p.grad.numpy()
        out[f"c{case}_loss"] = loss.detach().numpy()
        out[f"c{case}_items"] = items.numpy()
        print("loss case", case, float(loss), items.tolist())
    np.savez_compressed(os.path.join(HERE, "loss_golden.npz"), **out)


if __name__ == "__main__":
    ref = ref_import.load()
    make_nms(ref)
    make_loss(ref)
