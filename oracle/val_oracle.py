"""CPU oracle for the validation statistics. TEST INFRASTRUCTURE ONLY (see oracle/nms_oracle.py for the import rule).

numpy restatement of
  * scripts/utils/general.py:203-230, 324-358    clip_coords, scale_coords (torch branch, fp32)
  * scripts/utils/train_utils.py:294-333         YoloValidator.process_batch
  * scripts/utils/metrics.py:446-548             compute_ap, ap_per_class
Pinned against the reference's own functions (tests/test_oracle_val.py, build container) and against the committed
fixture tests/golden/val_golden.npz (outputs of the unmodified reference, generator tests/golden/make_golden_val.py).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from oracle.nms_oracle import box_iou, xywh2xyxy

IOUV = np.linspace(0.5, 0.95, 10, dtype=np.float64).astype(np.float32)  # train_utils.py: torch.linspace(0.5, 0.95, 10)


def scale_coords(img1_shape: Sequence[float], coords: np.ndarray, img0_shape: Sequence[float],
                 ratio_pad: Optional[Sequence] = None) -> np.ndarray:
    """general.py:324-358: network-input xyxy -> native-image xyxy, clipped; fp32 like a torch float tensor."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = ((img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2)
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    c = coords.astype(np.float32).copy()
    c[:, [0, 2]] -= np.float32(pad[0])
    c[:, [1, 3]] -= np.float32(pad[1])
    c[:, :4] /= np.float32(gain)
    w0, h0 = np.float32(img0_shape[1]), np.float32(img0_shape[0])
    c[:, 0] = np.clip(c[:, 0], 0, w0)
    c[:, 1] = np.clip(c[:, 1], 0, h0)
    c[:, 2] = np.clip(c[:, 2], 0, w0)
    c[:, 3] = np.clip(c[:, 3], 0, h0)
    return c


def process_batch(detections: np.ndarray, labels: np.ndarray, iouv: np.ndarray = IOUV) -> np.ndarray:
    """train_utils.py:294-333: detections (N, 6) xyxy conf cls, labels (M, 5) cls xyxy -> correct (N, niou) bool."""
    det = detections.astype(np.float32)
    lab = labels.astype(np.float32)
    correct = np.zeros((det.shape[0], iouv.shape[0]), dtype=bool)
    if det.shape[0] == 0 or lab.shape[0] == 0:
        return correct
    iou = box_iou(lab[:, 1:], det[:, :4])
    li, di = np.nonzero((iou >= iouv[0]) & (lab[:, 0:1] == det[None, :, 5]))
    if li.size:
        m = np.stack((li.astype(np.float32), di.astype(np.float32), iou[li, di]), 1)  # torch.cat promotes the indices to fp32
        if li.size > 1:
            m = m[m[:, 2].argsort()[::-1]]                       # best IoU first
            m = m[np.unique(m[:, 1], return_index=True)[1]]      # one pair per detection (now in detection order)
            m = m[np.unique(m[:, 0], return_index=True)[1]]      # one pair per label: the first in detection order (:324 is commented out)
        correct[m[:, 1].astype(np.int64)] = m[:, 2:3] >= iouv
    return correct


def compute_ap(recall: np.ndarray, precision: np.ndarray) -> Tuple[float, np.ndarray, np.ndarray]:
    """metrics.py:446-473: precision envelope + 101-point interpolated area."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([1.0], precision, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    x = np.linspace(0, 1, 101)
    y = np.interp(x, mrec, mpre)
    ap = float(np.sum((y[1:] + y[:-1]) * np.diff(x)) / 2.0)  # np.trapz
    return ap, mpre, mrec


def ap_per_class(tp: np.ndarray, conf: np.ndarray, pred_cls: np.ndarray, target_cls: np.ndarray):
    """metrics.py:476-548 without the plots: returns p, r, ap (nc, niou), f1, classes (int32)."""
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes = np.unique(target_cls)
    nc = classes.shape[0]
    px = np.linspace(0, 1, 1000)
    ap, p, r = np.zeros((nc, tp.shape[1])), np.zeros((nc, 1000)), np.zeros((nc, 1000))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        n_l, n_p = int((target_cls == c).sum()), int(sel.sum())
        if n_p == 0 or n_l == 0:
            continue
        fpc = (1 - tp[sel]).cumsum(0)
        tpc = tp[sel].cumsum(0)
        recall = tpc / (n_l + 1e-16)
        r[ci] = np.interp(-px, -conf[sel], recall[:, 0], left=0)
        precision = tpc / (tpc + fpc)
        p[ci] = np.interp(-px, -conf[sel], precision[:, 0], left=1)
        for j in range(tp.shape[1]):
            ap[ci, j] = compute_ap(recall[:, j], precision[:, j])[0]
    f1 = 2 * p * r / (p + r + 1e-16)
    i = f1.mean(0).argmax()
    return p[:, i], r[:, i], ap, f1[:, i], classes.astype("int32")


def synth_case(seed: int, n_det: int = 120, n_lab: int = 25, nc: int = 6, img: float = 640.0):
    """Detections scattered around the labels (so that every IoU regime occurs) plus clutter; returns (det, labels)."""
    rng = np.random.default_rng(seed)
    cxy = rng.uniform(0.1, 0.9, (n_lab, 2)) * img
    wh = np.exp(rng.uniform(np.log(0.04), np.log(0.4), (n_lab, 2))) * img
    lab = np.concatenate((rng.integers(0, nc, (n_lab, 1)).astype(np.float64), cxy - wh / 2, cxy + wh / 2), 1)
    src = rng.integers(0, n_lab, n_det)
    jit = rng.normal(0, 0.08, (n_det, 4)) * np.tile(wh[src], 2)
    box = lab[src, 1:] + jit
    box[:, 2:] = np.maximum(box[:, 2:], box[:, :2] + 1.0)
    cls = np.where(rng.random(n_det) < 0.8, lab[src, 0], rng.integers(0, nc, n_det))
    det = np.concatenate((box, rng.random((n_det, 1)), cls[:, None]), 1)
    det = det[np.argsort(-det[:, 4])]
    return det.astype(np.float32), lab.astype(np.float32)
