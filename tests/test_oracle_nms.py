"""Pins oracle/nms_oracle.py to the reference non_max_suppression / batched_nms: bit-exact on the committed
golden fixture (tests/golden/nms_golden.npz, outputs of the unmodified reference) and, when /root/reference is
present, against the reference itself on larger synthetic tensors."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import nms_oracle, ref_import

GOLD = os.path.join(os.path.dirname(__file__), "golden", "nms_golden.npz")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))


def test_nms_oracle_matches_golden():
    import make_golden as mg

    z = np.load(GOLD)
    pred = torch.from_numpy(z["pred"])
    for si, kw in enumerate(mg.NMS_SETTINGS):
        got = nms_oracle.non_max_suppression(pred, **kw)
        for i, g in enumerate(got):
            want = z[f"nms{si}_img{i}"]
            if kw.get("nms_type", "nms") in ("matrix_nms", "merge_nms"):  # exp / mm rounding (SLEEF / BLAS vs numpy)
                assert g.shape == want.shape and np.array_equal(g.numpy()[:, 5], want[:, 5]), (si, i)
                assert np.allclose(g.numpy(), want, rtol=1e-5, atol=1e-4), (si, i)
                continue
            assert g.shape == want.shape and np.array_equal(g.numpy(), want), (si, i)
    got = nms_oracle.batched_nms(pred, 0.05, 0.65, 100, False)
    for i, g in enumerate(got):
        assert np.array_equal(g.numpy(), z[f"bnms_img{i}"])


def test_xywh2xyxy_roundtrip_property():
    """tests/test_utils_general.py:16-47 of the reference: xyxy2xywh(xywh2xyxy(x)) == x."""
    x = torch.rand(100, 4).numpy() * 100 + 1
    y = nms_oracle.xywh2xyxy(x)
    back = np.stack(((y[:, 0] + y[:, 2]) / 2, (y[:, 1] + y[:, 3]) / 2, y[:, 2] - y[:, 0], y[:, 3] - y[:, 1]), 1)
    assert np.allclose(back, x, rtol=1e-5, atol=1e-4)


ALT = [dict(conf_thres=0.25, iou_thres=0.45, nms_type="fast_nms"), dict(conf_thres=0.3, iou_thres=0.5, nms_type="fast_nms", agnostic=True, multi_label=True),
       dict(conf_thres=0.25, iou_thres=0.45, nms_type="matrix_nms"), dict(conf_thres=0.25, iou_thres=0.45, nms_type="merge_nms"),
       dict(conf_thres=0.25, iou_thres=0.6, nms_type="merge_nms", multi_label=True, max_det=20)]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("kw", ALT)
def test_alt_nms_types_match_reference(kw):
    """fast / matrix / merge NMS: index selection identical, merged boxes and decayed scores within fp32 rounding
    (the reference's mm / exp run through BLAS / SLEEF, the oracle's through numpy)."""
    ref = ref_import.load()
    pred = nms_oracle.synth_predictions(2, n=2500, seed=7)
    a = ref.non_max_suppression(pred.clone(), **kw)
    b = nms_oracle.non_max_suppression(pred.clone(), **kw)
    for x, y in zip(a, b):
        assert x.shape == y.shape and x.shape[0] > 0
        assert torch.equal(x[:, 5], y[:, 5])
        assert torch.allclose(x, y, rtol=1e-5, atol=1e-4)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("kw", [dict(conf_thres=0.25, iou_thres=0.45), dict(conf_thres=0.25, iou_thres=0.45, multi_label=True),
                                dict(conf_thres=0.3, iou_thres=0.65, agnostic=True), dict(conf_thres=0.01, iou_thres=0.6, nms_type="batched_nms")])
def test_nms_oracle_matches_reference(kw):
    ref = ref_import.load()
    pred = nms_oracle.synth_predictions(2, n=3000, seed=4)
    a = ref.non_max_suppression(pred.clone(), **kw)
    b = nms_oracle.non_max_suppression(pred.clone(), **kw)
    for x, y in zip(a, b):
        assert x.shape == y.shape and torch.equal(x, y)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("nms_type", ["nms", "batched_nms", "fast_nms", "matrix_nms", "merge_nms"])
@pytest.mark.parametrize("agnostic", [False, True])
def test_val2_batched_nms_types_match_reference(nms_type, agnostic):
    """scripts/utils/nms.py:63-110 (the val2 path): every nms_type of the oracle against the unmodified reference --
    row selection identical, decayed scores / merged boxes within fp32 rounding."""
    ref = ref_import.load()
    pred = nms_oracle.synth_predictions(2, n=1500, nc=6, seed=11)
    a = ref.batched_nms(pred.clone(), conf_thres=0.2, iou_thres=0.6, nms_box=300, agnostic=agnostic, nms_type=nms_type)
    b = nms_oracle.batched_nms(pred.clone(), conf_thres=0.2, iou_thres=0.6, nms_box=300, agnostic=agnostic, nms_type=nms_type)
    for x, y in zip(a, b):
        x = x.float()
        assert x.shape == y.shape and x.shape[0] > 0
        assert torch.equal(x[:, 5], y[:, 5])
        if nms_type in ("matrix_nms", "merge_nms"):
            assert torch.allclose(x, y, rtol=1e-5, atol=1e-4)
        else:
            assert torch.equal(x, y)
