"""Validation statistics on the GPU (ay2_match_detections, ayolov2_b200.val_stats) vs the pinned CPU oracle
(oracle/val_oracle.py <- YoloValidator.process_batch / scale_coords / ap_per_class of the reference) and vs the committed
golden fixture produced by the unmodified reference. `correct` matrices must be identical (boolean)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "val_golden.npz")


def test_process_batch_matches_golden_and_oracle():
    from ayolov2_b200 import val_stats
    from oracle import val_oracle

    z = np.load(GOLD)
    for case in range(4):
        got = val_stats.process_batch(torch.from_numpy(z[f"c{case}_detn"]).cuda(), torch.from_numpy(z[f"c{case}_labn"]).cuda())
        assert np.array_equal(got.cpu().numpy(), z[f"c{case}_correct"]), case
    for seed in range(20, 26):
        det, lab = val_oracle.synth_case(seed, n_det=300, n_lab=60, nc=12)
        got = val_stats.process_batch(torch.from_numpy(det).cuda(), torch.from_numpy(lab).cuda())
        assert np.array_equal(got.cpu().numpy(), val_oracle.process_batch(det, lab))
    empty = val_stats.process_batch(torch.zeros((0, 6)).cuda(), torch.from_numpy(lab).cuda())
    assert empty.shape == (0, 10)
    none = val_stats.process_batch(torch.from_numpy(det).cuda(), torch.zeros((0, 5)).cuda())
    assert not none.any()


def test_batched_matching_with_native_mapping_and_ap():
    """Whole batch in one launch straight from an NMS-style buffer, labels in xywh network pixels, per-image letterbox
    meta; per image identical to oracle scale_coords + process_batch; the epoch reduction equals the oracle's ap_per_class."""
    from ayolov2_b200 import val_stats
    from oracle import nms_oracle, val_oracle

    B, max_det, nc = 5, 300, 8
    det = torch.zeros((B, max_det, 6))
    counts = torch.zeros(B, dtype=torch.int32)
    targets, shapes = [], []
    want_tp, want_conf, want_cls, want_t = [], [], [], []
    for b in range(B):
        d, l = val_oracle.synth_case(100 + b, n_det=40 * b, n_lab=12, nc=nc)  # image 0: no detections
        if b == 3:
            l = l[:0]  # image 3: no labels
        shape0 = (480 + 40 * b, 640 - 30 * b)
        ratio_pad = ((0.75 + 0.05 * b,) * 2, (8.0 * b, 12.0 + b))
        shapes.append((shape0, ratio_pad))
        det[b, :d.shape[0]] = torch.from_numpy(d)
        counts[b] = d.shape[0]
        xywh = np.stack(((l[:, 1] + l[:, 3]) / 2, (l[:, 2] + l[:, 4]) / 2, l[:, 3] - l[:, 1], l[:, 4] - l[:, 2]), 1).astype(np.float32)
        targets.append(np.concatenate((np.full((l.shape[0], 1), b, np.float32), l[:, :1], xywh), 1))
        dn = d.copy()
        if d.shape[0]:
            dn[:, :4] = val_oracle.scale_coords((640, 640), d[:, :4], shape0, ratio_pad)
        ln = np.concatenate((l[:, :1], val_oracle.scale_coords((640, 640), nms_oracle.xywh2xyxy(xywh), shape0, ratio_pad)), 1) \
            if l.shape[0] else np.zeros((0, 5), np.float32)
        want_tp.append(val_oracle.process_batch(dn, ln)); want_conf.append(d[:, 4]); want_cls.append(d[:, 5]); want_t.append(l[:, 0])
    targets = torch.from_numpy(np.concatenate(targets))
    vs = val_stats.ValStats(nc)
    correct = vs.update(det.cuda(), counts.cuda(), targets.cuda(), (640, 640), shapes).cpu().numpy().astype(bool)
    for b in range(B):
        assert np.array_equal(correct[b, :int(counts[b])], want_tp[b]), b
        assert not correct[b, int(counts[b]):].any()
    res = vs.compute()
    p, r, ap, f1, cls = val_oracle.ap_per_class(np.concatenate(want_tp), np.concatenate(want_conf), np.concatenate(want_cls),
                                                np.concatenate(want_t))
    assert np.array_equal(res["ap_class"], cls)
    assert np.allclose(res["ap50"], ap[:, 0], rtol=1e-12) and np.allclose(res["ap"], ap.mean(1), rtol=1e-12)
    assert np.allclose(res["p"], p, rtol=1e-12) and np.allclose(res["r"], r, rtol=1e-12)
    assert res["seen"] == B and res["nt"].sum() == targets.shape[0]
