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


class CandidateTable:
    """The candidate rows of a whole batch in the reference's order (ay2_nms_candidate_table): `rows` fp32 [B, cap, 8] =
    {x1, y1, x2, y2, conf, cls, 0, 0}, `counts` int32 [B], `max_coord` fp32 [B]. Built with one launch, no host sync; the
    capacity grows (and the max_nms re-ranking of metrics.py:378-379 is applied) only when the device flags ask for it,
    which the caller learns at the one synchronisation it needs anyway to size its output list."""

    def __init__(self, pred: torch.Tensor, conf_thres: float, multi_label: bool, classes: Optional[Sequence[int]],
                 max_nms: int, cap: Optional[int] = None):
        from . import _lib

        if not pred.is_cuda:
            raise RuntimeError("ayolov2_b200.nms runs on CUDA tensors only (no CPU fallback)")
        self.pred = pred.float().contiguous()
        B, n, no = self.pred.shape
        nc = no - 5
        self.B, self.n, self.no = B, n, no
        self.multi_label = bool(multi_label) and nc > 1
        self.max_nms = max_nms
        self.cmask = None
        if classes is not None:
            self.cmask = torch.zeros(nc, dtype=torch.uint8, device=pred.device)
            self.cmask[torch.as_tensor(list(classes), dtype=torch.long, device=pred.device)] = 1
        self.p = _lib.NmsParams()
        self.p.batch, self.p.n, self.p.no = B, n, no
        self.p.conf_thres, self.p.multi_label, self.p.max_nms = float(conf_thres), int(self.multi_label), int(max_nms)
        full = n * nc if self.multi_label else n
        self._build(min(full, cap if cap is not None else max(n, 4096)))
        self._full = full

    def _build(self, cap: int) -> None:
        from . import _lib

        dev = self.pred.device
        self.cap = max(int(cap), 1)
        self.rows = torch.empty((self.B, self.cap, 8), dtype=torch.float32, device=dev)
        self.counts = torch.empty(self.B, dtype=torch.int32, device=dev)
        self.max_coord = torch.empty(self.B, dtype=torch.float32, device=dev)
        self.flags = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.check(_lib.load().ay2_nms_candidate_table(self.pred.data_ptr(), C_byref(self.p), _lib.ptr(self.cmask),
                                                       self.rows.data_ptr(), self.cap, self.counts.data_ptr(),
                                                       self.max_coord.data_ptr(), self.flags.data_ptr(),
                                                       _lib.current_stream_ptr()), "ay2_nms_candidate_table")

    def settle(self) -> bool:
        """One host read of the flags. Returns True when the table had to be rebuilt (the caller re-runs its rule)."""
        f = int(self.flags.item())
        if not f:
            return False
        if f & 1:  # truncated: room for every (row, class) pair
            self._build(self._full)
            f = int(self.flags.item())
        if f & 2:  # more than max_nms candidates somewhere: keep each image's max_nms best, best first (stable)
            conf = self.rows[:, :, 4].clone()
            idx = torch.arange(self.cap, device=conf.device)[None, :]
            conf[idx >= self.counts[:, None]] = -float("inf")
            order = torch.sort(conf, dim=1, descending=True, stable=True).indices
            ranked = torch.gather(self.rows, 1, order[:, :, None].expand(-1, -1, 8))
            over = (self.counts > self.max_nms)[:, None, None]
            self.rows = torch.where(over, ranked, self.rows).contiguous()
            self.counts = torch.clamp(self.counts, max=self.max_nms)
            live = (idx < self.counts[:, None])[:, :, None]
            self.max_coord = torch.where(live, self.rows[:, :, :4], torch.full_like(self.rows[:, :, :4], -float("inf"))).amax((1, 2))
            self.flags.zero_()
        return True


def C_byref(x):
    import ctypes

    return ctypes.byref(x)


def _emit(out: torch.Tensor, count: torch.Tensor) -> List[torch.Tensor]:
    counts = count.tolist()
    return [out[i, :c].clone() for i, c in enumerate(counts)]


def _fast_or_matrix(tab: CandidateTable, rule: str, class_offset: float, iou_thres: float, out_cap: int) -> List[torch.Tensor]:
    from . import _lib

    lib = _lib.load()
    while True:
        dev = tab.rows.device
        ws = torch.empty((tab.B, tab.cap), dtype=torch.float32, device=dev)
        cap = max(1, min(out_cap, tab.cap))
        out = torch.zeros((tab.B, cap, 6), dtype=torch.float32, device=dev)
        cnt = torch.zeros(tab.B, dtype=torch.int32, device=dev)
        st = _lib.current_stream_ptr()
        if rule == "fast_nms":
            _lib.check(lib.ay2_nms_fast(tab.rows.data_ptr(), tab.counts.data_ptr(), tab.B, tab.cap, float(class_offset),
                                        float(iou_thres), ws.data_ptr(), out.data_ptr(), cap, cnt.data_ptr(), st), "ay2_nms_fast")
        else:
            _lib.check(lib.ay2_nms_matrix(tab.rows.data_ptr(), tab.counts.data_ptr(), tab.B, tab.cap, float(class_offset),
                                          ws.data_ptr(), out.data_ptr(), cap, cnt.data_ptr(), st), "ay2_nms_matrix")
        if not tab.settle():
            return _emit(out, cnt)


def _greedy(tab: CandidateTable, conf_thres: float, iou_thres: float, class_separated: bool, max_det: int,
            coordinate_trick: bool = False) -> ops.NmsWorkspace:
    """Greedy NMS of the table's batch on the batched kernel (ay2_nms_batched / ay2_nms_batched_scaled): survivors in kept
    order in `ws.out` / `ws.count`. The kernel draws the same candidates from the prediction tensor as the table holds."""
    from . import _lib

    B, n, no = tab.B, tab.n, tab.no
    ws = _workspace(B, n, no, max_det, tab.multi_label, tab.pred.device, max_candidates=tab._full if tab.multi_label else None)
    p = ws.p
    p.multi_label = int(tab.multi_label)
    p.conf_thres, p.iou_thres = float(conf_thres), float(iou_thres)
    p.agnostic, p.max_nms, p.max_wh = int(not class_separated), int(tab.max_nms), 4096.0
    fn, extra = ("ay2_nms_batched_scaled", (tab.max_coord.data_ptr(),)) if coordinate_trick else ("ay2_nms_batched", ())
    _lib.check(getattr(_lib.load(), fn)(tab.pred.data_ptr(), C_byref(p), _lib.ptr(tab.cmask), *extra, ws.ws.data_ptr(),
                                        ws.ws.numel(), ws.out.data_ptr(), ws.count.data_ptr(), ws.overflow.data_ptr(),
                                        _lib.current_stream_ptr()), fn)
    ws._keepalive = (tab,)
    return ws


def _merge(tab: CandidateTable, ws: ops.NmsWorkspace, class_offset: float, iou_thres: float, n_min_excl: int, n_max_excl: int) -> None:
    from . import _lib

    _lib.check(_lib.load().ay2_nms_merge(tab.rows.data_ptr(), tab.counts.data_ptr(), tab.B, tab.cap, float(class_offset),
                                         float(iou_thres), n_min_excl, n_max_excl, ws.out.data_ptr(), ws.out.shape[1],
                                         ws.count.data_ptr(), _lib.current_stream_ptr()), "ay2_nms_merge")


def _nms_other_types(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nms_type) -> list:
    """nms_type "batched_nms" / "fast_nms" / "matrix_nms" / "merge_nms" of non_max_suppression (metrics.py:388-431), the
    whole batch per launch (csrc/nms_variants.cu); the single host synchronisation is the read of the per-image counts."""
    tab = CandidateTable(prediction, conf_thres, multi_label, classes, max_nms=30000)
    sep = 0.0 if agnostic else 4096.0  # metrics.py:326 max_wh
    if nms_type == "fast_nms":
        return _fast_or_matrix(tab, nms_type, sep, iou_thres, max_det)
    if nms_type == "matrix_nms":
        return _fast_or_matrix(tab, nms_type, 0.0, iou_thres, max_det)  # metrics.py:405: the boxes are not class-separated
    while True:
        ws = _greedy(tab, conf_thres, iou_thres, class_separated=not agnostic, max_det=max_det,
                     coordinate_trick=nms_type == "batched_nms")
        if nms_type == "merge_nms":
            _merge(tab, ws, sep, iou_thres, 1, 3000)
        if not tab.settle():
            return _emit(ws.out, ws.count)


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
    """Reference signature of scripts/utils/nms.py:15-116 (the val2 path), every nms_type.

    The reference keeps the `nms_box` rows of highest objectness per image (:41-42), scores every (row, class) pair
    conf = cls * obj > conf_thres (:45-47) and then applies the chosen rule per image -- for "nms" / "merge_nms" class-
    separated (offset 4096 * class) only when `agnostic` is True (:58-62; the flag is inverted relative to
    non_max_suppression), for "fast_nms" / "matrix_nms" always class-separated (:75,84), for "batched_nms" through
    torchvision's coordinate trick. Here the gathered (B, nms_box, no) rows go through the same batched kernels as
    non_max_suppression (multi-label candidate rule, no max_det / max_nms cut). The reference has no bound on the number
    of survivors; images that fill the batched kernel's 1024-entry kept list are finished by the unbounded box-list NMS
    (ay2_nms_boxes), so nothing is truncated."""
    if nms_type not in ("nms", "batched_nms", "fast_nms", "matrix_nms", "merge_nms"):
        raise AssertionError("Wrong NMS type!!")
    if not prediction.is_cuda:
        raise RuntimeError("ayolov2_b200.nms runs on CUDA tensors only (no CPU fallback)")
    pred = prediction.float()
    B, n, no = pred.shape
    nc = no - 5
    k = min(nms_box, n)
    idx = pred[:, :, 4].argsort(descending=True)[:, :k]                         # nms.py:41 (same call, same tie behaviour)
    top = torch.gather(pred, 1, idx[:, :, None].expand(B, k, no)).contiguous()  # nms.py:42
    tab = CandidateTable(top, conf_thres, True, None, max_nms=k * max(nc, 1), cap=k * max(nc, 1))
    if nms_type in ("fast_nms", "matrix_nms"):
        return _fast_or_matrix(tab, nms_type, 4096.0, iou_thres, tab.cap)
    cap = 1024  # kept-list capacity of the batched kernel
    ws = _greedy(tab, conf_thres, iou_thres, class_separated=(agnostic or nms_type == "batched_nms"), max_det=cap,
                 coordinate_trick=nms_type == "batched_nms")
    full = (ws.count >= cap)
    if nms_type == "merge_nms":
        _merge(tab, ws, 4096.0 if agnostic else 0.0, iou_thres, -1, 1 << 30)
    counts = ws.count.tolist()
    out = [ws.out[i, :c].clone() for i, c in enumerate(counts)]
    for i in torch.nonzero(full).flatten().tolist():  # more than 1024 survivors: unbounded route for this image
        out[i] = _unbounded_image(tab, i, iou_thres, agnostic, nms_type)
    return out


def _unbounded_image(tab: CandidateTable, i: int, iou_thres: float, agnostic: bool, nms_type: str) -> torch.Tensor:
    """One image of the val2 path whose survivors exceed the batched kernel's kept list: box-list NMS without a bound."""
    n = int(tab.counts[i])
    x = tab.rows[i, :n, :6].clone()
    if nms_type == "batched_nms":
        boxes = x[:, :4] + (x[:, 5] * (x[:, :4].max() + 1))[:, None]
    else:
        boxes = x[:, :4] + x[:, 5:6] * 4096.0 if agnostic else x[:, :4]
    keep = nms_boxes(boxes.contiguous(), x[:, 4], iou_thres)
    if nms_type == "merge_nms":
        hit = box_iou(boxes[keep], boxes) > iou_thres
        w = hit * x[None, :, 4]
        x[keep, :4] = (w @ x[:, :4]) / w.sum(1, keepdim=True)
        keep = keep[hit.sum(1) > 1]
    return x[keep]
