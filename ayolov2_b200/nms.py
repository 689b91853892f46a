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
    if nms_type != "nms":
        raise NotImplementedError(f"nms_type={nms_type!r}: only the default 'nms' path runs on the B200 kernels")
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
