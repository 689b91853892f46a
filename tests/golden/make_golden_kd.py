"""Generates tests/golden/kd_golden.npz from the UNMODIFIED reference (build container only):
SoftTeacherTrainer.prepare_labels_for_augmention + the non-augmenting label assembly of get_pseudo_labeled_batch
(scripts/train/kd_trainer.py:385-397,417,436-487) on seeded synthetic NMS outputs (oracle.kd_oracle.synth_detections)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import kd_oracle, ref_import  # noqa: E402

CASES = [dict(seed=0, batch=6, image_size=(640, 640), thr=0.4, min_size=8.0),
         dict(seed=1, batch=4, image_size=(480, 480), thr=0.0, min_size=0.0),
         dict(seed=2, batch=5, image_size=(640, 640), thr=0.9, min_size=None)]


def reference_labels(kd, preds, image_size, thr, min_size):
    fake = types.SimpleNamespace(cfg_train={"image_size": image_size}, filter_invalid=kd.SoftTeacherTrainer.filter_invalid)
    labels_yolo = kd.SoftTeacherTrainer.prepare_labels_for_augmention(fake, [torch.from_numpy(p) for p in preds], thr=thr, min_size=min_size)
    rows = []
    for idx, cls_ids_bboxes in enumerate(labels_yolo):  # kd_trainer.py:388-397
        batch_ids = np.array([idx] * len(cls_ids_bboxes))
        rows.append(np.hstack([batch_ids[:, np.newaxis], cls_ids_bboxes]))
    return torch.Tensor(np.vstack(rows)).numpy()  # :417


def main():
    kd = ref_import.load_kd_trainer()
    out = {}
    for ci, c in enumerate(CASES):
        preds = kd_oracle.synth_detections(c["seed"], c["batch"], c["image_size"])
        out[f"c{ci}_labels"] = reference_labels(kd, preds, c["image_size"], c["thr"], c["min_size"])
        for i, p in enumerate(preds):
            out[f"c{ci}_pred{i}"] = p
    np.savez_compressed(os.path.join(HERE, "kd_golden.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("labels")})


if __name__ == "__main__":
    main()
