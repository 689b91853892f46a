"""Bit-exact parity of the batched CUDA NMS (ay2_nms_batched) with the CPU oracle (oracle/nms_oracle.py,
itself pinned to scripts/utils/metrics.py:285-443 of the reference)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cmp(pred, max_det=300, **kw):
    from ayolov2_b200.nms import non_max_suppression
    from oracle import nms_oracle

    want = nms_oracle.non_max_suppression(pred, max_det=max_det, **kw)
    got = non_max_suppression(pred.cuda(), max_det=max_det, **kw)
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape, (i, g.shape, w.shape)
        assert torch.equal(g.cpu(), w), f"image {i}: first diff row {(g.cpu() != w).any(1).nonzero()[:3].tolist()}"


@pytest.mark.parametrize("kw", [
    dict(conf_thres=0.25, iou_thres=0.45),
    dict(conf_thres=0.25, iou_thres=0.45, multi_label=True),
    dict(conf_thres=0.3, iou_thres=0.65, agnostic=True),
    dict(conf_thres=0.25, iou_thres=0.45, classes=[1, 5, 7]),
    dict(conf_thres=0.6, iou_thres=0.3),
], ids=["default", "multilabel", "agnostic", "classes", "sparse"])
def test_nms_synthetic(kw):
    from oracle import nms_oracle

    _cmp(nms_oracle.synth_predictions(4, n=25200, seed=0), **kw)


def test_nms_small_and_empty():
    from oracle import nms_oracle

    pred = nms_oracle.synth_predictions(3, n=1000, seed=3)
    pred[1, :, 4] = 0.0  # image with no candidates
    _cmp(pred, conf_thres=0.25, iou_thres=0.45)
    _cmp(pred, conf_thres=0.25, iou_thres=0.45, max_det=17)
    _cmp(pred, conf_thres=0.001, iou_thres=0.65, multi_label=True)  # > 8192 candidates -> global-memory sort path


def test_nms_ties_documented():
    """Equal scores: the lower candidate index wins (torchvision's stable descending sort)."""
    pred = torch.zeros(1, 8, 7)
    pred[0, :, :4] = torch.tensor([100., 100., 50., 50.])
    pred[0, :, 4] = 0.9
    pred[0, :, 5] = 0.8
    pred[0, 4:, 0] += 200.0
    _cmp(pred, conf_thres=0.25, iou_thres=0.45)


def test_nms_paths_agree_wide_boxes_and_single_class():
    """The class-partitioned fast path must hand over to the general chunked scan when a box leaves the max_wh window
    (cross-class overlap possible) or when one class dominates; both must equal the oracle."""
    from oracle import nms_oracle

    pred = nms_oracle.synth_predictions(2, n=3000, seed=5)
    pred[0, :50, 2:4] = 6000.0  # huge boxes: classes overlap despite the offset
    _cmp(pred, conf_thres=0.25, iou_thres=0.45)
    one = nms_oracle.synth_predictions(2, n=3000, nc=1, seed=6)
    _cmp(one, conf_thres=0.25, iou_thres=0.45)
    many = nms_oracle.synth_predictions(1, n=6000, seed=7, cand_frac=0.9)  # > 4096 candidates: general path
    _cmp(many, conf_thres=0.25, iou_thres=0.45)
