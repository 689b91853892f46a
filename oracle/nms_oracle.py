"""CPU oracle for the NMS post-process. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product path (ayolov2_b200/) never does.

A plain numpy/torch-CPU restatement of
  * scripts/utils/metrics.py:285-443   non_max_suppression (every nms_type: "nms", "batched_nms", "fast_nms",
                                       "matrix_nms", "merge_nms")
  * scripts/utils/metrics.py:138-164   box_iou
  * scripts/utils/general.py:297-321   xywh2xyxy
  * scripts/utils/nms.py:15-116        batched_nms (val2 path, every nms_type)
  * torchvision.ops.nms (torchvision 0.10.1 pinned by environment.yml:28; 0.26 behaves the same): stable
    descending score sort, greedy suppression with fp32 IoU = inter / (area_i + area_j - inter), strict `>`.

Pinned against the reference's own functions (imported from /root/reference in the build container) by
tests/test_oracle_nms.py and against the committed fixtures in tests/golden/nms_*.npz.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    """general.py:316-319 with the default ratio/wh/pad (1*1*(x -/+ w/2) + 0), fp32."""
    x = x.astype(np.float32, copy=False)
    y = np.empty_like(x)
    one = np.float32(1.0)
    zero = np.float32(0.0)
    half_w = x[:, 2] / np.float32(2)
    half_h = x[:, 3] / np.float32(2)
    y[:, 0] = one * one * (x[:, 0] - half_w) + zero
    y[:, 1] = one * one * (x[:, 1] - half_h) + zero
    y[:, 2] = one * one * (x[:, 0] + half_w) + zero
    y[:, 3] = one * one * (x[:, 1] + half_h) + zero
    return y


def greedy_nms(boxes: np.ndarray, scores: np.ndarray, iou_thres: float, limit: Optional[int] = None) -> np.ndarray:
    """torchvision.ops.nms on CPU: returns kept indices in descending-score order.

    `limit` stops after that many kept boxes (the reference slices i[:max_det] afterwards, metrics.py:386-387;
    stopping early returns the same prefix because greedy decisions never depend on later boxes)."""
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    boxes = boxes.astype(np.float32, copy=False)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep: List[int] = []
    thr = float(iou_thres)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(int(i))
        if limit is not None and len(keep) >= limit:
            break
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def box_iou(box1: np.ndarray, box2: np.ndarray) -> np.ndarray:
    """metrics.py:138-164: (N, M) IoU of xyxy boxes, fp32, inter / (area1 + area2 - inter) (0/0 -> nan like the reference)."""
    b1, b2 = box1.astype(np.float32, copy=False), box2.astype(np.float32, copy=False)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = np.minimum(b1[:, None, 2:], b2[None, :, 2:]) - np.maximum(b1[:, None, :2], b2[None, :, :2])
    wh = np.maximum(wh, np.float32(0))
    inter = wh[..., 0] * wh[..., 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / (a1[:, None] + a2[None, :] - inter)).astype(np.float32)


def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        classes: Optional[Sequence[int]] = None, agnostic: bool = False, multi_label: bool = False,
                        max_det: int = 300, nms_type: str = "nms") -> List[torch.Tensor]:
    """metrics.py:285-443 restated (all five nms_type values); prediction fp32 [B, n, 5+nc] on CPU."""
    pred = prediction.detach().cpu().float().numpy()
    nc = pred.shape[2] - 5
    max_wh = np.float32(4096)  # metrics.py:326
    max_nms = 30000  # metrics.py:327
    conf_t = np.float32(conf_thres)
    multi_label = bool(multi_label) and nc > 1  # :330
    out: List[torch.Tensor] = []
    for xi in range(pred.shape[0]):
        x = pred[xi]
        x = x[x[:, 4] > conf_t].copy()  # :313,337
        if not x.shape[0]:
            out.append(torch.zeros((0, 6)))
            continue
        x[:, 5:] *= x[:, 4:5]  # :353
        box = xywh2xyxy(x[:, :4])  # :356
        if multi_label:  # :359-361
            i, j = np.nonzero(x[:, 5:] > conf_t)
            x = np.concatenate((box[i], x[i, j + 5, None], j[:, None].astype(np.float32)), 1)
        else:  # :362-364
            j = np.argmax(x[:, 5:], axis=1)
            conf = x[np.arange(x.shape[0]), j + 5]
            x = np.concatenate((box, conf[:, None], j[:, None].astype(np.float32)), 1)[conf > conf_t]
        if classes is not None:  # :367-368
            x = x[np.isin(x[:, 5], np.asarray(classes, dtype=np.float32))]
        n = x.shape[0]
        if not n:
            out.append(torch.zeros((0, 6)))
            continue
        if n > max_nms:  # :378-379
            x = x[np.argsort(-x[:, 4], kind="stable")[:max_nms]]
        if nms_type == "nms":
            c = x[:, 5:6] * (np.float32(0) if agnostic else max_wh)  # :383
            boxes = x[:, :4] + c
        elif nms_type == "batched_nms":  # :391-394 -> torchvision _batched_nms_coordinate_trick
            cls = x[:, 5] * 0 if agnostic else x[:, 5]
            max_coord = x[:, :4].max()
            boxes = x[:, :4] + (cls * (max_coord + np.float32(1)))[:, None]
        elif nms_type == "fast_nms":  # :397-401 (yolact): a box survives when no EARLIER row overlaps it (rows keep x's order)
            cls = x[:, 5] * 0 if agnostic else x[:, 5]
            boxes = x[:, :4] + cls[:, None] * max_wh
            iou = np.triu(box_iou(boxes, boxes), k=1)
            keep = np.nonzero(iou.max(0) < np.float32(iou_thres))[0]
            out.append(torch.from_numpy(x[keep][:max_det].astype(np.float32)))
            continue
        elif nms_type == "matrix_nms":  # :404-413: gaussian score decay (sigma 0.5), nothing is removed, scores change in place
            iou = np.triu(box_iou(x[:, :4], x[:, :4]), k=1)
            m = iou.max(0)[:, None]
            decay = torch.exp(torch.from_numpy(-(iou ** 2 - m ** 2) / np.float32(0.5))).numpy().min(0)
            x = x.copy()
            x[:, 4] *= decay
            out.append(torch.from_numpy(x[:max_det].astype(np.float32)))
            continue
        elif nms_type == "merge_nms":  # :414-431: greedy NMS, then every kept box <- score-weighted mean of its overlaps
            c = x[:, 5:6] * (np.float32(0) if agnostic else max_wh)
            boxes = (x[:, :4] + c).astype(np.float32)
            keep = greedy_nms(boxes, x[:, 4], iou_thres)[:max_det]
            if 1 < n < 3e3:
                hit = box_iou(boxes[keep], boxes) > np.float32(iou_thres)
                weights = hit * x[None, :, 4]
                x = x.copy()
                x[keep, :4] = (weights @ x[:, :4]).astype(np.float32) / weights.sum(1, keepdims=True)
                keep = keep[hit.sum(1) > 1]  # redundant=True (:326): require at least one other overlapping box
            out.append(torch.from_numpy(x[keep].astype(np.float32)))
            continue
        else:
            raise NotImplementedError(nms_type)
        keep = greedy_nms(boxes.astype(np.float32), x[:, 4], iou_thres, limit=max_det)
        out.append(torch.from_numpy(x[keep[:max_det]].astype(np.float32)))
    return out


def batched_nms(prediction: torch.Tensor, conf_thres: float = 0.001, iou_thres: float = 0.65, nms_box: int = 500,
                agnostic: bool = False, nms_type: str = "nms") -> List[torch.Tensor]:
    """nms.py:15-116 restated, every nms_type (note: for "nms" / "merge_nms" class offsets are applied only when
    `agnostic`, :58-62; "fast_nms" / "matrix_nms" always offset by 4096 * class, :75,84)."""
    pred = prediction.detach().cpu().float().numpy()
    out: List[torch.Tensor] = []
    conf_t = np.float32(conf_thres)
    thr32 = np.float32(iou_thres)
    for xi in range(pred.shape[0]):
        x = pred[xi]
        idx = np.argsort(-x[:, 4], kind="stable")[:nms_box]  # :41
        o = x[idx]
        confs = o[:, 5:] * o[:, 4:5]  # :45
        j, k = np.nonzero(confs > conf_t)  # :46
        xywh = o[j, :4]
        two = np.float32(2.0)
        box = np.stack((xywh[:, 0] - xywh[:, 2] / two, xywh[:, 1] - xywh[:, 3] / two, xywh[:, 0] + xywh[:, 2] / two,
                        xywh[:, 1] + xywh[:, 3] / two), 1).astype(np.float32)  # :50-54
        det = np.concatenate((box, confs[j, k, None], k[:, None].astype(np.float32)), 1).astype(np.float32)
        if agnostic:  # :58-60 (sic)
            bboxes = (det[:, :4] + det[:, 5:6] * np.float32(4096)).astype(np.float32)
        else:
            bboxes = det[:, :4]
        if nms_type == "nms":  # :64-65
            keep = greedy_nms(bboxes, det[:, 4], iou_thres)
        elif nms_type == "batched_nms":  # :68-72 -> torchvision coordinate trick on the un-offset boxes
            if det.shape[0] == 0:
                keep = np.zeros((0,), dtype=np.int64)
            else:
                off = det[:, 5] * (det[:, :4].max() + np.float32(1))
                keep = greedy_nms((det[:, :4] + off[:, None]).astype(np.float32), det[:, 4], iou_thres)
        elif nms_type in ("fast_nms", "matrix_nms"):  # :75-99
            if det.shape[0] == 0:  # `continue` at :80 / :91 leaves the (empty) candidate rows in place
                out.append(torch.from_numpy(det))
                continue
            sep = (det[:, :4] + det[:, 5:6] * np.float32(4096)).astype(np.float32)
            iou = np.triu(box_iou(sep, sep), k=1)
            if nms_type == "fast_nms":
                keep = np.nonzero(iou.max(0) < thr32)[0]
            else:
                m = iou.max(0)[:, None]
                decay = torch.exp(torch.from_numpy(-(iou ** 2 - m ** 2) / np.float32(0.5))).numpy().min(0)
                det = det.copy()
                det[:, 4] *= decay
                keep = np.arange(det.shape[0])
        elif nms_type == "merge_nms":  # :101-110
            keep = greedy_nms(bboxes, det[:, 4], iou_thres)
            if det.shape[0]:
                hit = box_iou(bboxes[keep], bboxes) > thr32
                weights = hit * det[None, :, 4]
                det = det.copy()
                det[keep, :4] = (weights @ det[:, :4]).astype(np.float32) / weights.sum(1, keepdims=True)
                keep = keep[hit.sum(1) > 1]
        else:
            raise NotImplementedError(nms_type)
        out.append(torch.from_numpy(det[keep].astype(np.float32)))
    return out


def synth_predictions(batch: int, n: int = 25200, nc: int = 80, seed: int = 0, cand_frac: float = 0.08,
                      clusters: int = 200, img: float = 640.0) -> torch.Tensor:
    """Synthetic (batch, n, 5+nc) prediction tensor that forces real suppression chains:
    ~cand_frac of the rows have objectness in (0.25, 1), boxes are jittered copies of `clusters` centres."""
    g = torch.Generator().manual_seed(seed)
    pred = torch.zeros(batch, n, 5 + nc)
    cls = torch.rand(batch, n, nc, generator=g)
    boost = torch.randint(0, nc, (batch, n), generator=g)
    cls.scatter_(2, boost[..., None], 0.5 + 0.5 * torch.rand(batch, n, 1, generator=g))
    obj = torch.rand(batch, n, generator=g) * 0.2
    is_c = torch.rand(batch, n, generator=g) < cand_frac
    obj = torch.where(is_c, 0.25 + 0.75 * torch.rand(batch, n, generator=g), obj)
    centres = torch.rand(batch, clusters, 2, generator=g) * img
    which = torch.randint(0, clusters, (batch, n), generator=g)
    xy = torch.gather(centres, 1, which[..., None].expand(-1, -1, 2)) + 4.0 * torch.randn(batch, n, 2, generator=g)
    lo, hi = np.log(16.0), np.log(256.0)
    cwh = torch.exp(lo + (hi - lo) * torch.rand(batch, clusters, 2, generator=g))
    wh = torch.gather(cwh, 1, which[..., None].expand(-1, -1, 2)) * torch.exp(0.1 * torch.randn(batch, n, 2, generator=g))
    # same-cluster boxes share a class most of the time so that suppression actually happens
    ccls = torch.randint(0, nc, (batch, clusters), generator=g)
    same = torch.rand(batch, n, generator=g) < 0.8
    forced = torch.gather(ccls, 1, which)
    cls2 = cls.clone()
    cls2.scatter_(2, forced[..., None], 0.9 + 0.1 * torch.rand(batch, n, 1, generator=g))
    cls = torch.where(same[..., None], cls2, cls)
    pred[..., :2] = xy
    pred[..., 2:4] = wh
    pred[..., 4] = obj
    pred[..., 5:] = cls
    return pred.contiguous()
