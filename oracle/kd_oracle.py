"""CPU oracle for the knowledge-distillation pseudo-labels. TEST INFRASTRUCTURE ONLY (see oracle/nms_oracle.py for the import rule).

numpy restatement of
  * scripts/train/kd_trainer.py:465-487   SoftTeacherTrainer.filter_invalid
  * scripts/train/kd_trainer.py:436-463   prepare_labels_for_augmention (normalise, clip, xyxy -> xywh)
  * scripts/utils/general.py:250-295      xyxy2xywh with check_validity (wh = (1, 1), clip_eps None)
  * scripts/train/kd_trainer.py:385-397   the non-augmenting label assembly of get_pseudo_labeled_batch
Pinned against the UNMODIFIED reference functions imported from /root/reference (tests/test_oracle_kd.py, build container) and
against tests/golden/kd_golden.npz (outputs of the unmodified reference, generator tests/golden/make_golden_kd.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def filter_invalid(pred: np.ndarray, thr: float = 0.0, min_size: Optional[float] = 0.0) -> np.ndarray:
    """kd_trainer.py:465-487 on one image's (n, 6) detections; fp32 comparisons like a torch float tensor against a scalar."""
    pred = np.asarray(pred, np.float32).reshape(-1, 6)
    pred = pred[pred[:, 4] > np.float32(thr)]
    if min_size is not None:
        bw, bh = pred[:, 2] - pred[:, 0], pred[:, 3] - pred[:, 1]
        pred = pred[(bw > np.float32(min_size)) & (bh > np.float32(min_size))]
    return pred


def xyxy2xywh_valid(x: np.ndarray) -> np.ndarray:
    """general.py:276-295 for normalised fp32 boxes (wh = (1, 1)), validity correction on."""
    x = np.asarray(x, np.float32)
    y = np.copy(x)
    y[:, 0] = (x[:, 0] + x[:, 2]) / 2
    y[:, 1] = (x[:, 1] + x[:, 3]) / 2
    y[:, 2] = x[:, 2] - x[:, 0]
    y[:, 3] = x[:, 3] - x[:, 1]
    y[:, 2] = y[:, 2] + (np.minimum((y[:, 0] - (y[:, 2] / 2)), 0) * 2)
    y[:, 2] = y[:, 2] - ((np.maximum((y[:, 0] + (y[:, 2] / 2)), 1) - 1) * 2)
    y[:, 3] = y[:, 3] + (np.minimum((y[:, 1] - (y[:, 3] / 2)), 0) * 2)
    y[:, 3] = y[:, 3] - ((np.maximum((y[:, 1] + (y[:, 3] / 2)), 1) - 1) * 2)
    return y.clip(1e-12, 1)


def prepare_labels(preds: Sequence[np.ndarray], image_size: Sequence[int], thr: float = 0.0,
                   min_size: Optional[float] = 0.0) -> List[np.ndarray]:
    """kd_trainer.py:436-463: per image (n_i, 5) = class, x, y, w, h."""
    width, height = image_size
    whwh = np.array([width, height, width, height])
    out = []
    for pred in preds:
        p = filter_invalid(pred, thr, min_size)
        if len(p) == 0:
            out.append(np.zeros((0, 5)))
            continue
        boxes = p[:, :4].copy()
        boxes /= whwh  # float32 /= int64: computed in double, rounded once
        boxes.clip(min=0, max=1, out=boxes)
        out.append(np.hstack([p[:, 5][:, np.newaxis], xyxy2xywh_valid(boxes)]))
    return out


def pseudo_labels(preds: Sequence[np.ndarray], image_size: Sequence[int], thr: float = 0.0,
                  min_size: Optional[float] = 0.0) -> np.ndarray:
    """kd_trainer.py:385-397 + :417: (N, 6) fp32 = image, class, x, y, w, h (what `torch.Tensor(...)` of the stack holds)."""
    rows = []
    for idx, l in enumerate(prepare_labels(preds, image_size, thr, min_size)):
        rows.append(np.hstack([np.array([idx] * len(l))[:, np.newaxis], l]))
    return np.vstack(rows).astype(np.float32) if rows else np.zeros((0, 6), np.float32)


def synth_detections(seed: int, batch: int, image_size: Sequence[int], max_n: int = 40) -> List[np.ndarray]:
    """NMS-like outputs: boxes partly outside the image, tiny boxes, scores over the whole range, an empty image."""
    rng = np.random.default_rng(seed)
    w, h = image_size
    out = []
    for b in range(batch):
        n = 0 if b == 1 else int(rng.integers(1, max_n))
        cx, cy = rng.uniform(-20, w + 20, n), rng.uniform(-20, h + 20, n)
        bw, bh = np.exp(rng.uniform(np.log(0.5), np.log(w), n)), np.exp(rng.uniform(np.log(0.5), np.log(h), n))
        det = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2, rng.uniform(0, 1, n), rng.integers(0, 80, n)], 1)
        out.append(det.astype(np.float32))
    return out
