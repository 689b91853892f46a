"""Drop-in replacements for the reference NMS entry points, running on the batched CUDA kernels.

  non_max_suppression  <-  scripts/utils/metrics.py:285-443
  batched_nms          <-  scripts/utils/nms.py:15-116

Same names, argument meaning and return convention (list of (n_i, 6) tensors [x1, y1, x2, y2, conf, cls]).
The whole batch is processed by ay2_nms_batched in a fixed number of launches; the only host
synchronisation is the final read-back of the per-image counts that the list-of-tensors API implies.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import ops

_WS_CACHE: Dict[Tuple, ops.NmsWorkspace] = {}


def _workspace(batch: int, n: int, no: int, max_det: int, multi_label: bool, device, max_candidates=None):
    key = (batch, n, no, max_det, multi_label, str(device), max_candidates)
    ws = _WS_CACHE.get(key)
    if ws is None:
        if len(_WS_CACHE) > 8:
            _WS_CACHE.clear()
        ws = ops.NmsWorkspace(batch, n, no, max_det=max_det, multi_label=multi_label, max_candidates=max_candidates,
                              device=device)
        _WS_CACHE[key] = ws
    return ws


def nms_device(prediction: torch.Tensor, conf_thres: float, iou_thres: float, agnostic: bool = False,
               multi_label: bool = False, max_det: int = 300, classes: Optional[Sequence[int]] = None,
               workspace: Optional[ops.NmsWorkspace] = None) -> ops.NmsWorkspace:
    """Asynchronous batched NMS: returns the workspace whose `.out` [B, max_det, 6] / `.count` [B] hold the result
    (valid once the current stream reaches this point; no host sync)."""
    if not prediction.is_cuda:
        raise RuntimeError("ayolov2_b200.nms runs on CUDA tensors only (no CPU fallback)")
    pred = prediction
    if pred.dtype != torch.float32:
        pred = pred.float()
    pred = pred.contiguous()
    B, n, no = pred.shape
    nc = no - 5
    multi_label = bool(multi_label) and nc > 1  # metrics.py:330
    ws = workspace or _workspace(B, n, no, max_det, multi_label, pred.device)
    cmask = None
    if classes is not None:
        cmask = torch.zeros(nc, dtype=torch.uint8, device=pred.device)
        cmask[torch.as_tensor(list(classes), dtype=torch.long, device=pred.device)] = 1
    ws.p.multi_label = int(multi_label)
    ws.run(pred, conf_thres, iou_thres, agnostic=agnostic, class_mask=cmask)
    ws._keepalive = (pred, cmask)
    return ws


def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        classes: Optional[list] = None, agnostic: bool = False, multi_label: bool = False,
                        labels: Union[tuple, list] = (), max_det: int = 300, nms_type: str = "nms") -> list:
    """Run NMS on inference results (reference signature, metrics.py:285-295)."""
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if nms_type not in ("nms", "batched_nms", "fast_nms", "matrix_nms", "merge_nms"):
        raise AssertionError("Wrong NMS type!!")  # metrics.py:433
    if labels:
        # metrics.py:340-346: a-priori labels are appended as rows with obj = 1 and a one-hot class
        nc = prediction.shape[2] - 5
        extra = max(len(l) for l in labels)
        if extra:
            pad = torch.zeros((prediction.shape[0], extra, nc + 5), device=prediction.device, dtype=prediction.dtype)
            for xi, lab in enumerate(labels):
                if len(lab):
                    pad[xi, :len(lab), :4] = lab[:, 1:5]
                    pad[xi, :len(lab), 4] = 1.0
                    pad[xi, range(len(lab)), lab[:, 0].long() + 5] = 1.0
            prediction = torch.cat((prediction, pad), 1)
    if nms_type != "nms":
        return _nms_other_types(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nms_type)
    ws = nms_device(prediction, conf_thres, iou_thres, agnostic=agnostic, multi_label=multi_label, max_det=max_det,
                    classes=classes)
    counts = ws.count.tolist()  # the one host sync
    if int(ws.overflow.item()):
        # more candidates than the list capacity: redo with room for every (row, class) pair
        B, n, no = prediction.shape
        big = ops.NmsWorkspace(B, n, no, max_det=max_det, multi_label=bool(ws.p.multi_label),
                               max_candidates=n * (no - 5), device=prediction.device)
        ws = nms_device(prediction, conf_thres, iou_thres, agnostic=agnostic, multi_label=multi_label, max_det=max_det,
                        classes=classes, workspace=big)
        counts = ws.count.tolist()
    out = ws.out
    return [out[i, :c].clone() for i, c in enumerate(counts)]


def nms_boxes(boxes: torch.Tensor, scores: torch.Tensor, iou_thres: float) -> torch.Tensor:
    """torchvision.ops.nms(boxes, scores, iou_thres) on the ay2_nms_boxes kernels: kept indices in descending score order
    (stable for ties), greedy suppression IoU > iou_thres. One host sync (the count)."""
    from . import _lib

    if not boxes.is_cuda:
        raise RuntimeError("ayolov2_b200.nms.nms_boxes runs on CUDA tensors only (no CPU fallback)")
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros((0,), dtype=torch.long, device=boxes.device)
    b = boxes.float().contiguous()
    order = torch.sort(scores.float(), descending=True, stable=True).indices.to(torch.int32)
    mask = torch.empty((n * ((n + 63) // 64),), dtype=torch.int64, device=b.device)
    keep = torch.empty((n,), dtype=torch.int32, device=b.device)
    count = torch.zeros((1,), dtype=torch.int32, device=b.device)
    _lib.check(_lib.load().ay2_nms_boxes(b.data_ptr(), order.data_ptr(), n, float(iou_thres), mask.data_ptr(), keep.data_ptr(),
                                         count.data_ptr(), _lib.current_stream_ptr()), "ay2_nms_boxes")
    return keep[:int(count.item())].long()


def _xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    """general.py:316-319 with the default ratio / wh / pad, same fp32 operation order."""
    y = torch.empty_like(x)
    hw, hh = x[:, 2] / 2, x[:, 3] / 2
    y[:, 0], y[:, 1], y[:, 2], y[:, 3] = x[:, 0] - hw, x[:, 1] - hh, x[:, 0] + hw, x[:, 1] + hh
    return y


def _nms_other_types(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nms_type) -> list:
    """The non-default nms_type branches of metrics.py:388-431 ("batched_nms", "fast_nms", "matrix_nms", "merge_nms"):
    per image, candidate rows in the reference's order (metrics.py:337-379), then the branch's matrix arithmetic on the
    CUDA IoU matrix (ay2_box_iou) / box-list NMS (ay2_nms_boxes). Device tensors throughout; these are diagnostic
    variants in the reference (the validator and val.py use "nms"), so they are not batched across images."""
    if not prediction.is_cuda:
        raise RuntimeError("ayolov2_b200.nms runs on CUDA tensors only (no CPU fallback)")
    pred = prediction.float()
    nc = pred.shape[2] - 5
    max_wh, max_nms = 4096.0, 30000  # metrics.py:326-327
    multi_label = bool(multi_label) and nc > 1
    out = [torch.zeros((0, 6), device=pred.device)] * pred.shape[0]
    for xi in range(pred.shape[0]):
        x = pred[xi]
        x = x[x[:, 4] > conf_thres].clone()                                       # :313,337
        if not x.shape[0]:
            continue
        x[:, 5:] *= x[:, 4:5]                                                     # :353
        box = _xywh2xyxy(x[:, :4])                                                # :356
        if multi_label:                                                           # :359-361
            i, j = (x[:, 5:] > conf_thres).nonzero(as_tuple=False).T
            x = torch.cat((box[i], x[i, j + 5, None], j[:, None].float()), 1)
        else:                                                                     # :362-364
            conf, j = x[:, 5:].max(1, keepdim=True)
            x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        if classes is not None:                                                   # :367-368
            x = x[(x[:, 5:6] == torch.tensor(classes, device=x.device)).any(1)]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:                                                           # :378-379
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        if nms_type == "batched_nms":      # :391-394 (torchvision's coordinate trick: offset = class * (max coordinate + 1))
            c = x[:, 5] * 0 if agnostic else x[:, 5]
            boxes = x[:, :4] + (c * (x[:, :4].max() + 1))[:, None]
            out[xi] = x[nms_boxes(boxes, x[:, 4], iou_thres)[:max_det]]
        elif nms_type == "fast_nms":       # :397-401
            c = x[:, 5] * 0 if agnostic else x[:, 5]
            boxes = x[:, :4] + c.view(-1, 1) * max_wh
            iou = box_iou(boxes, boxes).triu_(diagonal=1)
            out[xi] = x[iou.max(0)[0] < iou_thres][:max_det]
        elif nms_type == "matrix_nms":     # :404-413
            iou = box_iou(x[:, :4], x[:, :4]).triu_(diagonal=1)
            m = iou.max(0)[0].view(-1, 1)
            x[:, 4] *= torch.exp(-(iou ** 2 - m ** 2) / 0.5).min(0)[0]
            out[xi] = x[:max_det]
        else:                              # merge_nms, :414-431
            boxes, scores = x[:, :4] + x[:, 5:6] * (0 if agnostic else max_wh), x[:, 4]
            i = nms_boxes(boxes, scores, iou_thres)[:max_det]
            if 1 < n < 3e3:
                hit = box_iou(boxes[i], boxes) > iou_thres
                weights = hit * scores[None]
                x[i, :4] = torch.mm(weights, x[:, :4]).float() / weights.sum(1, keepdim=True)
                i = i[hit.sum(1) > 1]      # redundant = True (:326)
            out[xi] = x[i]
    return out


def box_iou(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
    """(N, M) pairwise IoU of xyxy boxes (reference signature, metrics.py:138-164), one CUDA launch."""
    from . import _lib

    if not (box1.is_cuda and box2.is_cuda):
        raise RuntimeError("ayolov2_b200.nms.box_iou runs on CUDA tensors only (no CPU fallback)")
    b1, b2 = box1.float().contiguous(), box2.float().contiguous()
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    _lib.check(_lib.load().ay2_box_iou(b1.data_ptr(), b1.shape[0], b2.data_ptr(), b2.shape[0], out.data_ptr(),
                                       _lib.current_stream_ptr()), "ay2_box_iou")
    return out


def batched_nms(prediction: torch.Tensor, conf_thres: float = 0.001, iou_thres: float = 0.65, nms_box: int = 500,
                agnostic: bool = False, nms_type: str = "nms") -> List[torch.Tensor]:
    """Reference signature of scripts/utils/nms.py:15-116 (the val2.py path), nms_type "nms".

    The reference keeps the `nms_box` rows of highest objectness per image (:41-42), scores every (row, class) pair
    conf = cls * obj > conf_thres (:45-47) and runs torchvision NMS per image -- class-separated (offset 4096 * class) only
    when `agnostic` is True, plain otherwise (:58-62; the flag is inverted relative to non_max_suppression). That is the
    multi-label candidate generation + greedy suppression of ay2_nms_batched on the gathered (B, nms_box, no) rows with
    the agnostic flag flipped and no max_det / max_nms cut, so the batch runs in the same fixed number of launches.
    Candidate order (row in objectness order, then class) is the reference's, which fixes torchvision's stable tie order.
    """
    if nms_type != "nms":
        raise NotImplementedError(f"nms_type={nms_type!r}: only 'nms' runs on the B200 kernels (box_iou is available for the others)")
    if not prediction.is_cuda:
        raise RuntimeError("ayolov2_b200.nms runs on CUDA tensors only (no CPU fallback)")
    pred = prediction.float()
    B, n, no = pred.shape
    nc = no - 5
    k = min(nms_box, n)
    idx = pred[:, :, 4].argsort(descending=True)[:, :k]                      # nms.py:41 (same call, same tie behaviour)
    top = torch.gather(pred, 1, idx[:, :, None].expand(B, k, no)).contiguous()  # nms.py:42
    cap = 1024  # kMaxDetCap of the kernel
    ws = _workspace(B, k, no, cap, True, pred.device, max_candidates=k * nc)
    ws.p.multi_label = 1
    ws.run(top, conf_thres, iou_thres, agnostic=not agnostic, max_nms=k * nc)
    ws._keepalive = (top,)
    counts = ws.count.tolist()
    if max(counts, default=0) >= cap:
        raise NotImplementedError(f"an image kept >= {cap} boxes; raise conf_thres (the kernel's survivor list holds {cap})")
    return [ws.out[i, :c].clone() for i, c in enumerate(counts)]
