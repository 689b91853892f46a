"""Pins oracle/val_oracle.py (scale_coords, process_batch, ap_per_class restatements) to the reference: bit-exact on the
committed fixture tests/golden/val_golden.npz (outputs of the unmodified reference) and, when /root/reference is present,
against the reference's own functions on fresh seeds."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_import, val_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "val_golden.npz")
RATIO_PAD, SHAPE0 = ((0.8, 0.8), (16.0, 24.0)), (740, 760)


def test_val_oracle_matches_golden():
    z = np.load(GOLD)
    tps, confs, pcls, tcls = [], [], [], []
    for case in range(4):
        det, lab = z[f"c{case}_det"], z[f"c{case}_lab"]
        dn = det.copy()
        dn[:, :4] = val_oracle.scale_coords((640, 640), det[:, :4], SHAPE0, RATIO_PAD)
        ln = lab.copy()
        ln[:, 1:] = val_oracle.scale_coords((640, 640), lab[:, 1:], SHAPE0, RATIO_PAD)
        assert np.array_equal(dn, z[f"c{case}_detn"]) and np.array_equal(ln, z[f"c{case}_labn"])
        correct = val_oracle.process_batch(dn, ln)
        assert np.array_equal(correct, z[f"c{case}_correct"]), case
        tps.append(correct); confs.append(det[:, 4]); pcls.append(det[:, 5]); tcls.append(lab[:, 0])
    p, r, ap, f1, cls = val_oracle.ap_per_class(np.concatenate(tps), np.concatenate(confs), np.concatenate(pcls), np.concatenate(tcls))
    assert np.array_equal(cls, z["ap_cls"])
    for got, key in ((p, "ap_p"), (r, "ap_r"), (ap, "ap_ap"), (f1, "ap_f1")):
        assert np.allclose(got, z[key], rtol=1e-12, atol=1e-15), key


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_val_oracle_matches_reference_fresh_seeds():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_val as mg

    process_batch, scale_coords, ap_per_class = mg.ref_functions()
    tps, confs, pcls, tcls = [], [], [], []
    for seed in range(10, 16):
        det, lab = val_oracle.synth_case(seed, n_det=300, n_lab=40, nc=10)
        a = process_batch(torch.from_numpy(det), torch.from_numpy(lab)).numpy()
        b = val_oracle.process_batch(det, lab)
        assert np.array_equal(a, b) and a.any()
        dn = torch.from_numpy(det[:, :4].copy())
        scale_coords((640, 640), dn, (480, 600))  # ratio_pad computed from the shapes
        assert np.array_equal(dn.numpy(), val_oracle.scale_coords((640, 640), det[:, :4], (480, 600)))
        tps.append(b); confs.append(det[:, 4]); pcls.append(det[:, 5]); tcls.append(lab[:, 0])
    ra = ap_per_class(np.concatenate(tps), np.concatenate(confs), np.concatenate(pcls), np.concatenate(tcls))
    rb = val_oracle.ap_per_class(np.concatenate(tps), np.concatenate(confs), np.concatenate(pcls), np.concatenate(tcls))
    for x, y in zip(ra, rb):
        assert np.allclose(x, y, rtol=1e-12, atol=1e-15)


def test_product_host_reductions_match_oracle():
    """ayolov2_b200.val_stats (host side, no GPU): ap_per_class / compute_ap / scale_meta against the pinned oracle."""
    from ayolov2_b200 import val_stats

    tps, confs, pcls, tcls = [], [], [], []
    for seed in range(30, 35):
        det, lab = val_oracle.synth_case(seed, n_det=250, n_lab=35, nc=9)
        tps.append(val_oracle.process_batch(det, lab)); confs.append(det[:, 4]); pcls.append(det[:, 5]); tcls.append(lab[:, 0])
    args = (np.concatenate(tps), np.concatenate(confs), np.concatenate(pcls), np.concatenate(tcls))
    for x, y in zip(val_stats.ap_per_class(*args), val_oracle.ap_per_class(*args)):
        assert np.allclose(x, y, rtol=1e-12, atol=1e-15)
    meta = val_stats.scale_meta((640, 640), [((480, 600), ((0.8, 0.8), (16.0, 24.0))), (480, 600)], device="cpu").numpy()
    assert np.allclose(meta[0], [0.8, 16.0, 24.0, 600.0, 480.0])
    gain = min(640 / 480, 640 / 600)  # general.py:343-349
    assert np.allclose(meta[1], [gain, (640 - 600 * gain) / 2, (640 - 480 * gain) / 2, 600.0, 480.0])
    with pytest.raises(RuntimeError):
        val_stats.match_batch(torch.zeros((1, 4, 6)), torch.zeros(1, dtype=torch.int32), torch.zeros((0, 6)), torch.linspace(0.5, 0.95, 10))
