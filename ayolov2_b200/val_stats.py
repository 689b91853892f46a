"""Validation statistics behind the reference's validator API, on the GPU.

  process_batch / match_batch  <-  YoloValidator.process_batch + statistics_per_image (scripts/utils/train_utils.py:294-401)
  scale_coords meta            <-  scripts/utils/general.py:324-358
  ap_per_class / compute_ap    <-  scripts/utils/metrics.py:446-548 (end-of-epoch numpy on the host in the reference too)

The reference walks the NMS output image by image and moves every image's IoU matches to the host
(`.cpu().numpy()`, train_utils.py:319-329). Here the whole batch is matched by ONE launch (ay2_match_detections) straight
from the NMS output buffer; the "correct" matrices stay on the device until `ValStats.compute()` reads them back once.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_MAX_LABELS_SMEM = 4096  # labels per image the matcher stages in shared memory


def iou_thresholds(device="cuda") -> torch.Tensor:
    """The validator's `iouv` (10 thresholds 0.5:0.95)."""
    return torch.linspace(0.5, 0.95, 10, device=device)


def scale_meta(img_hw: Tuple[int, int], shapes: Sequence, device="cuda") -> torch.Tensor:
    """Per-image {gain, pad_x, pad_y, native_w, native_h} of scale_coords (general.py:342-357). `shapes[i]` is the data
    loader's `((h0, w0), ((ratio_h, ratio_w), (pad_w, pad_h)))` tuple, or just `(h0, w0)` (gain / pad derived)."""
    rows = []
    for s in shapes:
        if len(s) == 2 and not isinstance(s[0], (tuple, list, np.ndarray, torch.Tensor)):
            shape0, ratio_pad = s, None
        else:
            shape0, ratio_pad = s[0], (s[1] if len(s) > 1 else None)
        h0, w0 = float(shape0[0]), float(shape0[1])
        if ratio_pad is None:
            gain = min(img_hw[0] / h0, img_hw[1] / w0)
            pad = ((img_hw[1] - w0 * gain) / 2, (img_hw[0] - h0 * gain) / 2)
        else:
            gain, pad = float(ratio_pad[0][0]), (float(ratio_pad[1][0]), float(ratio_pad[1][1]))
        rows.append([gain, pad[0], pad[1], w0, h0])
    return torch.tensor(rows, dtype=torch.float32, device=device)


def match_batch(det: torch.Tensor, counts: torch.Tensor, labels: torch.Tensor, iouv: torch.Tensor,
                meta: Optional[torch.Tensor] = None, labels_cap: Optional[int] = None) -> torch.Tensor:
    """det [B, max_det, 6] + counts [B] (the NMS output buffers), labels [T, 6] = (image, class, box) -> uint8
    [B, max_det, niou]. meta None: label boxes are xyxy in the detections' coordinates; else xywh network-input pixels and
    both sides are mapped to the native image (see scale_meta). No host synchronisation."""
    if not det.is_cuda:
        raise RuntimeError("ayolov2_b200.val_stats runs on CUDA tensors only (no CPU fallback)")
    det = det.float().contiguous()
    counts = counts.to(torch.int32).contiguous()
    labels = labels.to(det.device).float().contiguous()
    iouv = iouv.to(det.device).float().contiguous()
    B, max_det, _ = det.shape
    nt = labels.shape[0]
    if labels_cap is None:
        labels_cap = nt
        if nt > _MAX_LABELS_SMEM:  # many labels: size the staging buffer by the busiest image (one host sync)
            labels_cap = int(torch.bincount(labels[:, 0].long(), minlength=B).max().item())
    correct = torch.empty((B, max_det, iouv.numel()), dtype=torch.uint8, device=det.device)
    _lib.check(_lib.load().ay2_match_detections(det.data_ptr(), counts.data_ptr(), B, max_det, labels.data_ptr() if nt else None, nt,
                                                max(labels_cap, 1), _lib.ptr(meta), iouv.data_ptr(), iouv.numel(),
                                                correct.data_ptr(), _lib.current_stream_ptr()), "ay2_match_detections")
    return correct


def process_batch(detections: torch.Tensor, labels: torch.Tensor, iouv: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Reference signature (train_utils.py:294-333): detections (N, 6) xyxy conf cls, labels (M, 5) cls xyxy ->
    correct (N, niou) bool, on the device."""
    iouv = iou_thresholds(detections.device) if iouv is None else iouv
    n = detections.shape[0]
    if n == 0:
        return torch.zeros((0, iouv.numel()), dtype=torch.bool, device=detections.device)
    lab6 = torch.cat((torch.zeros((labels.shape[0], 1), device=detections.device), labels.to(detections.device).float()), 1)
    counts = torch.tensor([n], dtype=torch.int32, device=detections.device)
    return match_batch(detections[None], counts, lab6, iouv)[0].bool()


def compute_ap(recall: np.ndarray, precision: np.ndarray) -> Tuple[float, np.ndarray, np.ndarray]:
    """metrics.py:446-473: monotone precision envelope, 101-point interpolation, trapezoid area."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.flip(np.maximum.accumulate(np.flip(np.concatenate(([1.0], precision, [0.0])))))
    x = np.linspace(0, 1, 101)
    y = np.interp(x, mrec, mpre)
    return float(((y[1:] + y[:-1]) * (x[1:] - x[:-1])).sum() / 2.0), mpre, mrec


def ap_per_class(tp: np.ndarray, conf: np.ndarray, pred_cls: np.ndarray, target_cls: np.ndarray):
    """metrics.py:476-548 (no plots): precision / recall at the best mean-F1 confidence, AP per class and IoU threshold."""
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes = np.unique(target_cls)
    grid = np.linspace(0, 1, 1000)
    ap = np.zeros((classes.shape[0], tp.shape[1]))
    p, r = np.zeros((classes.shape[0], 1000)), np.zeros((classes.shape[0], 1000))
    for ci, c in enumerate(classes):
        mine = pred_cls == c
        n_lab = int((target_cls == c).sum())
        if not mine.any() or n_lab == 0:
            continue
        tpc = tp[mine].cumsum(0)
        fpc = (1 - tp[mine]).cumsum(0)
        recall = tpc / (n_lab + 1e-16)
        precision = tpc / (tpc + fpc)
        r[ci] = np.interp(-grid, -conf[mine], recall[:, 0], left=0)
        p[ci] = np.interp(-grid, -conf[mine], precision[:, 0], left=1)
        ap[ci] = [compute_ap(recall[:, j], precision[:, j])[0] for j in range(tp.shape[1])]
    f1 = 2 * p * r / (p + r + 1e-16)
    best = f1.mean(0).argmax()
    return p[:, best], r[:, best], ap, f1[:, best], classes.astype("int32")


class ValStats:
    """Accumulates (correct, conf, pred class, target class) over validation batches on the device and reduces them
    like YoloValidator.compute_statistics (train_utils.py:475-512)."""

    def __init__(self, nc: int, device="cuda") -> None:
        self.nc = nc
        self.device = torch.device(device)
        self.iouv = iou_thresholds(self.device)
        self._chunks: List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = []
        self._tcls: List[torch.Tensor] = []
        self.seen = 0

    def update(self, det: torch.Tensor, counts: torch.Tensor, targets: torch.Tensor, img_hw: Tuple[int, int],
               shapes: Optional[Sequence] = None) -> torch.Tensor:
        """det / counts: NMS output of the batch; targets: (T, 6) image, class, xywh in network-input pixels (the validator
        has already multiplied by width / height, train_utils.py:120-123); shapes: the loader's per-image shape tuples (None:
        evaluate in network-input coordinates). Returns the batch's `correct` matrix (device)."""
        meta = scale_meta(img_hw, shapes, self.device) if shapes is not None else None
        labels = targets.to(self.device).float()
        if meta is None and labels.shape[0]:  # no native mapping: the matcher wants xyxy
            xy, half = labels[:, 2:4], labels[:, 4:6] / 2
            labels = torch.cat((labels[:, :2], xy - half, xy + half), 1)
        correct = match_batch(det, counts, labels, self.iouv, meta)
        self._chunks.append((correct, det[..., 4:6].clone(), counts.clone()))
        self._tcls.append(labels[:, 1].clone())
        self.seen += det.shape[0]
        return correct

    def compute(self) -> dict:
        tps, confs, pcls = [], [], []
        for correct, cc, counts in self._chunks:  # one read-back per batch tensor, at the end of the epoch
            cnt = counts.tolist()
            c_h, cc_h = correct.cpu().numpy(), cc.cpu().numpy()
            for i, n in enumerate(cnt):
                tps.append(c_h[i, :n].astype(bool))
                confs.append(cc_h[i, :n, 0])
                pcls.append(cc_h[i, :n, 1])
        tcls = torch.cat(self._tcls).cpu().numpy() if self._tcls else np.zeros((0,), np.float32)
        out = {"seen": self.seen, "nt": np.bincount(tcls.astype(np.int64), minlength=self.nc)}
        if tps and np.concatenate(tps).any():
            p, r, ap, f1, ap_class = ap_per_class(np.concatenate(tps), np.concatenate(confs), np.concatenate(pcls), tcls)
            out.update(p=p, r=r, f1=f1, ap_class=ap_class, ap50=ap[:, 0], ap=ap.mean(1), mp=p.mean(), mr=r.mean(),
                       map50=ap[:, 0].mean(), map=ap.mean())
        return out
