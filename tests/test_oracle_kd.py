"""Pins oracle/kd_oracle.py (KD pseudo-labels, kd_trainer.py:385-397,436-487; general.py:250-295) against the committed
outputs of the unmodified reference (tests/golden/kd_golden.npz) and against the reference itself when /root/reference is
present. CPU only."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

from make_golden_kd import CASES, reference_labels  # noqa: E402
from oracle import kd_oracle, ref_import  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "kd_golden.npz")


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_pseudo_labels_match_golden(ci):
    g, c = np.load(GOLD), CASES[ci]
    preds = [g[f"c{ci}_pred{i}"] for i in range(c["batch"])]
    for a, b in zip(preds, kd_oracle.synth_detections(c["seed"], c["batch"], c["image_size"])):
        assert np.array_equal(a, b)
    got = kd_oracle.pseudo_labels(preds, c["image_size"], c["thr"], c["min_size"])
    want = g[f"c{ci}_labels"]
    assert got.dtype == np.float32 and got.shape == want.shape and np.array_equal(got, want)
    assert np.all((got[:, 2:] >= np.float32(1e-12)) & (got[:, 2:] <= 1))


def test_validity_correction_keeps_boxes_inside_the_unit_square():
    preds = kd_oracle.synth_detections(9, 8, (640, 640), max_n=200)
    lab = kd_oracle.pseudo_labels(preds, (640, 640), 0.0, 0.0)
    assert len(lab) > 100
    assert np.all(lab[:, 2] - lab[:, 4] / 2 >= -1e-6) and np.all(lab[:, 2] + lab[:, 4] / 2 <= 1 + 1e-6)
    assert np.all(lab[:, 3] - lab[:, 5] / 2 >= -1e-6) and np.all(lab[:, 3] + lab[:, 5] / 2 <= 1 + 1e-6)
    assert np.all(np.diff(lab[:, 0]) >= 0)  # image order


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_oracle_matches_the_unmodified_reference():
    kd = ref_import.load_kd_trainer()
    for seed in range(6):
        size = [(640, 640), (416, 416), (320, 320)][seed % 3]
        thr, ms = [(0.3, 4.0), (0.0, 0.0), (0.7, None)][seed % 3]
        preds = kd_oracle.synth_detections(50 + seed, 7, size, max_n=120)
        want = reference_labels(kd, preds, size, thr, ms)
        got = kd_oracle.pseudo_labels(preds, size, thr, ms)
        assert got.shape == want.shape and np.array_equal(got, want)
