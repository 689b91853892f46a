"""KD pseudo-labels on the GPU (csrc/pseudo_labels.cu through the C-ABI, ayolov2_b200/kd.py) against the oracle
(oracle/kd_oracle.py, pinned to the unmodified reference) and the reference's committed outputs. fp32 arithmetic in the
reference's operation order: the bar is bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

pytestmark = pytest.mark.gpu

from make_golden_kd import CASES  # noqa: E402
from ayolov2_b200 import kd  # noqa: E402
from oracle import kd_oracle  # noqa: E402
from _parity import record  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "kd_golden.npz")


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_pseudo_labels_equal_reference_golden(ci):
    g, c = np.load(GOLD), CASES[ci]
    preds = [torch.from_numpy(g[f"c{ci}_pred{i}"]).cuda() for i in range(c["batch"])]
    per_image = kd.prepare_labels_for_augmention(preds, c["image_size"], c["thr"], c["min_size"])
    want = g[f"c{ci}_labels"]
    got = np.concatenate([np.concatenate([np.full((len(l), 1), i, np.float32), l.cpu().numpy()], 1) for i, l in enumerate(per_image)], 0)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_full_size_buffer_equals_oracle():
    """64 images x 300 detection slots (BASELINE batch, max_det): several 1024-slot chunks, ragged counts, empty images."""
    rng = np.random.default_rng(4)
    B, max_det = 64, 300
    preds = kd_oracle.synth_detections(77, B, (640, 640), max_n=max_det + 1)
    preds[5] = preds[5][:0]
    det = torch.zeros((B, max_det, 6), dtype=torch.float32)
    det[:] = torch.from_numpy(rng.uniform(0, 640, (B, max_det, 6)).astype(np.float32))  # garbage beyond the counts must be ignored
    for i, p in enumerate(preds):
        det[i, :len(p)] = torch.from_numpy(p)
    counts = torch.tensor([len(p) for p in preds], dtype=torch.int32)
    for thr, ms in ((0.25, 2.0), (0.0, None), (0.999, 0.0)):
        labels, per_image = kd.pseudo_labels_from_detections(det.cuda(), counts.cuda(), (640, 640), thr, ms)
        want = kd_oracle.pseudo_labels(preds, (640, 640), thr, ms)
        assert labels.shape == want.shape
        record(f"kd/pseudo_labels_64x300_thr{thr}_min{ms}", labels=int(len(want)),
               max_abs_diff=float(np.abs(labels.cpu().numpy() - want).max()) if len(want) else 0.0)
        assert np.array_equal(labels.cpu().numpy(), want)
        assert per_image.tolist() == [int((want[:, 0] == i).sum()) for i in range(B)]


def test_teacher_pipeline_feeds_compute_loss():
    """get_pseudo_labeled_batch: teacher forward + NMS + pseudo-labels on the device == the oracle applied to the same NMS
    output; the labels are a valid ComputeLoss target tensor."""
    from ayolov2_b200 import synth
    from ayolov2_b200.nms import non_max_suppression

    teacher = synth.build_model("yolov5n", seed=3).cuda().eval()
    B, H, W = 3, 256, 256
    img = torch.randint(0, 256, (B, 3, H, W), generator=torch.Generator().manual_seed(1), dtype=torch.uint8)
    sample = img.cuda().float() / 255.0
    synth.calibrate_head(teacher, lambda: teacher(sample)[1], cand_frac=0.1)
    imgs, labels = kd.get_pseudo_labeled_batch(teacher, img, (W, H), 0.25, 0.45, conf_thr=0.3, bbox_size_thr=4.0)
    dets = non_max_suppression(teacher(imgs)[0], 0.25, 0.45)
    want = kd_oracle.pseudo_labels([d.cpu().numpy() for d in dets], (W, H), 0.3, 4.0)
    assert len(want) > 5 and np.array_equal(labels.cpu().numpy(), want)
    assert imgs.dtype == torch.float32 and float(imgs.max()) <= 1.0
