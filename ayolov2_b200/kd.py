"""Knowledge-distillation caller of the NMS kernels: teacher detections -> pseudo-labels on the device.

The reference (scripts/train/kd_trainer.py:356-433) runs the teacher, `non_max_suppression`, then walks the per-image
detection lists on the host (`prepare_labels_for_augmention` :436-463, `filter_invalid` :465-487, numpy `xyxy2xywh`,
hstack / vstack, :385-397) before the labels go back to the device for `ComputeLoss`. Here the NMS output buffer never
leaves the device: `ay2_pseudo_labels` (csrc/pseudo_labels.cu) filters, normalises, converts and compacts the whole batch
in one launch into the (N, 6) tensor `ComputeLoss` consumes. No CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib


def pseudo_labels_from_detections(det: torch.Tensor, counts: torch.Tensor, image_size: Sequence[int], thr: float = 0.0,
                                  min_size: Optional[float] = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """det [B, max_det, 6] fp32 + counts [B] int32 (the NMS output buffers, e.g. `Detector.run_device`) ->
    (labels [N, 6] fp32 = image, class, x, y, w, h; per-image label counts [B] int32). `image_size` = (width, height) like
    `cfg_train["image_size"]` (kd_trainer.py:439). Synchronises once to learn N (the reference's `.cpu()` per image)."""
    if not det.is_cuda:
        raise RuntimeError("pseudo_labels_from_detections: CUDA tensors only (no CPU fallback)")
    assert det.dtype == torch.float32 and det.dim() == 3 and det.shape[2] == 6 and det.is_contiguous()
    B, max_det = det.shape[0], det.shape[1]
    counts = counts.to(device=det.device, dtype=torch.int32).contiguous()
    labels = torch.empty((B * max_det, 6), dtype=torch.float32, device=det.device)
    out_counts = torch.empty(B + 1, dtype=torch.int32, device=det.device)
    width, height = image_size
    _lib.check(_lib.load().ay2_pseudo_labels(det.data_ptr(), counts.data_ptr(), B, max_det, float(thr),
                                             float(min_size if min_size is not None else 0.0), int(min_size is not None),
                                             float(width), float(height), labels.data_ptr(), out_counts.data_ptr(),
                                             _lib.current_stream_ptr()), "ay2_pseudo_labels")
    n = int(out_counts[B].item())
    return labels[:n], out_counts[:B]


def _stack_detections(preds: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    dev = preds[0].device
    max_det = max(1, max(int(p.shape[0]) for p in preds))
    det = torch.zeros((len(preds), max_det, 6), dtype=torch.float32, device=dev)
    for i, p in enumerate(preds):
        det[i, :p.shape[0]] = p
    return det, torch.tensor([int(p.shape[0]) for p in preds], dtype=torch.int32, device=dev)


def prepare_labels_for_augmention(preds: List[torch.Tensor], image_size: Sequence[int], thr: float = 0.0,
                                  min_size: Optional[float] = 0.0) -> List[torch.Tensor]:
    """Same contract as SoftTeacherTrainer.prepare_labels_for_augmention (kd_trainer.py:436-463) on the list
    `non_max_suppression` returns: per image an (n_i, 5) tensor = class, x, y, w, h (device tensors instead of numpy arrays)."""
    if not preds:
        return []
    det, counts = _stack_detections(preds)
    labels, per_image = pseudo_labels_from_detections(det, counts, image_size, thr, min_size)
    return [l[:, 1:] for l in torch.split(labels, per_image.tolist())]


@torch.no_grad()
def get_pseudo_labeled_batch(teacher, imgs: torch.Tensor, image_size: Sequence[int], nms_conf_thr: float, nms_iou_thr: float,
                             conf_thr: float = 0.0, bbox_size_thr: Optional[float] = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """The non-augmenting branch of SoftTeacherTrainer.get_pseudo_labeled_batch (kd_trainer.py:374-397,417): teacher forward on
    `imgs / 255`, NMS, pseudo-labels. `teacher` is a kindle YOLOModel (this repo's) in eval mode; returns (imgs float, labels)."""
    from .nms import nms_device

    imgs = imgs.to(next(teacher.parameters()).device).float() / 255.0
    pred, _ = teacher(imgs)
    ws = nms_device(pred, nms_conf_thr, nms_iou_thr)
    labels, _ = pseudo_labels_from_detections(ws.out, ws.count, image_size, conf_thr, bbox_size_thr)
    return imgs, labels
