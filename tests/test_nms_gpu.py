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


def test_nms_segmented_rounds_and_fallbacks():
    """The segmented bit-matrix path: few classes -> long segments resolved over several lazy rounds; agnostic ->
    one segment; a segment whose triangular masks exceed the shared-memory budget hands over to the chunked scan."""
    from oracle import nms_oracle

    few = nms_oracle.synth_predictions(3, n=2400, nc=3, seed=11, cand_frac=0.7, clusters=12)  # ~1700 candidates, 3 classes
    _cmp(few, conf_thres=0.25, iou_thres=0.45)
    _cmp(few, conf_thres=0.25, iou_thres=0.6, max_det=50)
    agn = nms_oracle.synth_predictions(2, n=1200, seed=12, cand_frac=0.75, clusters=30)  # ~900 candidates, one segment
    _cmp(agn, conf_thres=0.25, iou_thres=0.45, agnostic=True)
    big = nms_oracle.synth_predictions(2, n=2400, nc=1, seed=13, cand_frac=0.8, clusters=40)  # one ~1900-row segment
    _cmp(big, conf_thres=0.25, iou_thres=0.45)


def test_nms_iou_threshold_sliver():
    """Pairs whose IoU sits within a few ulp of the threshold take the exact-division branch of the division-free
    filter: nested boxes with IoU = area ratio ~ iou_thres, swept in 1-ulp steps around it."""
    import numpy as np

    thr = 0.45
    n = 384  # 768 boxes, one segment (positions leave the max_wh window): 9.6k mask words, the bit-matrix path
    pred = torch.zeros(1, 2 * n, 6)
    base = torch.tensor([320.0, 320.0])
    w = 200.0
    for i in range(n):
        # outer box (score high) and an inner box whose area ratio is thr * (1 + (i - n/2) * 2^-24)
        ratio = np.float64(thr) * (1.0 + (i - n // 2) * 2.0 ** -24)
        iw = np.float32(w * np.sqrt(ratio))
        c = base + torch.tensor([float(i % 32) * 1300.0, float(i // 32) * 1300.0])  # far apart: pairs do not interact
        pred[0, 2 * i, :4] = torch.tensor([c[0], c[1], w, w])
        pred[0, 2 * i + 1, :4] = torch.tensor([c[0], c[1], float(iw), float(iw)])
        pred[0, 2 * i, 4], pred[0, 2 * i + 1, 4] = 0.9, 0.8
    pred[0, :, 5] = 1.0
    _cmp(pred, max_det=1000, conf_thres=0.25, iou_thres=thr)
    _cmp(pred, max_det=1000, conf_thres=0.25, iou_thres=float(np.float32(thr)))
    _cmp(pred[:, :256], conf_thres=0.25, iou_thres=thr, agnostic=True)


@pytest.mark.parametrize("agnostic,conf", [(False, 0.3), (True, 0.93), (False, 0.6)])
def test_batched_nms_val2_path(agnostic, conf):
    """scripts/utils/nms.py:15-116 (val2.py): top-nms_box rows by objectness, every (row, class) above conf, torchvision NMS
    with the (inverted) agnostic flag -- bit-exact against the oracle restatement pinned to the reference."""
    from ayolov2_b200.nms import batched_nms
    from oracle import nms_oracle

    pred = nms_oracle.synth_predictions(3, n=6000, nc=80, seed=11, cand_frac=0.15)
    want = nms_oracle.batched_nms(pred, conf_thres=conf, iou_thres=0.65, nms_box=500, agnostic=agnostic)
    got = batched_nms(pred.cuda(), conf_thres=conf, iou_thres=0.65, nms_box=500, agnostic=agnostic)
    assert sum(w.shape[0] for w in want) > 100 and max(w.shape[0] for w in want) < 1024
    for g, w in zip(got, want):
        assert g.shape == w.shape, (g.shape, w.shape)
        assert torch.equal(g.cpu(), w)


def test_box_iou_bit_exact():
    """scripts/utils/metrics.py:138-164 on identical fp32 inputs (the expression below is the reference's)."""
    from ayolov2_b200.nms import box_iou

    g = torch.Generator().manual_seed(4)

    def boxes(n):
        xy = torch.rand(n, 2, generator=g) * 600
        wh = torch.rand(n, 2, generator=g) * 200 + 1
        return torch.cat((xy, xy + wh), 1)

    b1, b2 = boxes(333), boxes(1025)
    area1 = (b1.T[2] - b1.T[0]) * (b1.T[3] - b1.T[1])
    area2 = (b2.T[2] - b2.T[0]) * (b2.T[3] - b2.T[1])
    inter = (torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])).clamp(0).prod(2)
    want = inter / (area1[:, None] + area2 - inter)
    got = box_iou(b1.cuda(), b2.cuda()).cpu()
    assert torch.equal(got, want)
    assert box_iou(b1[:0].cuda(), b2.cuda()).shape == (0, 1025)


@pytest.mark.parametrize("kw", [
    dict(conf_thres=0.05, iou_thres=0.6, nms_type="batched_nms"),
    dict(conf_thres=0.25, iou_thres=0.45, nms_type="batched_nms", agnostic=True, multi_label=True),
    dict(conf_thres=0.25, iou_thres=0.45, nms_type="fast_nms"),
    dict(conf_thres=0.3, iou_thres=0.5, nms_type="fast_nms", agnostic=True, multi_label=True),
    dict(conf_thres=0.25, iou_thres=0.45, nms_type="matrix_nms"),
    dict(conf_thres=0.25, iou_thres=0.45, nms_type="merge_nms"),
    dict(conf_thres=0.25, iou_thres=0.6, nms_type="merge_nms", multi_label=True, max_det=20),
], ids=["batched", "batched-agnostic-ml", "fast", "fast-agnostic-ml", "matrix", "merge", "merge-ml"])
def test_other_nms_types(kw):
    """metrics.py:388-431: index selection (which rows, which order, which class) identical to the pinned oracle; the
    decayed scores (exp) and merged boxes (mm) within fp32 rounding, tolerance 1e-5 relative / 1e-4 absolute pixels."""
    from ayolov2_b200.nms import non_max_suppression
    from oracle import nms_oracle

    pred = nms_oracle.synth_predictions(2, n=2500, seed=7)
    pred[1, :, 4] *= 0.5
    want = nms_oracle.non_max_suppression(pred, **kw)
    got = non_max_suppression(pred.cuda(), **kw)
    for g, w in zip(got, want):
        assert g.shape == w.shape and w.shape[0] > 0
        assert torch.equal(g[:, 5].cpu(), w[:, 5])
        if kw["nms_type"] in ("batched_nms", "fast_nms"):
            assert torch.equal(g.cpu(), w)
        else:
            assert torch.allclose(g.cpu(), w, rtol=1e-5, atol=1e-4)


def test_nms_boxes_matches_torchvision_semantics():
    """ay2_nms_boxes == the oracle's greedy_nms (torchvision.ops.nms restatement) on clustered boxes incl. exact ties."""
    from ayolov2_b200.nms import nms_boxes
    from oracle import nms_oracle

    g = torch.Generator().manual_seed(3)
    c = torch.rand(40, 2, generator=g) * 600
    xy = c[torch.randint(0, 40, (3000,), generator=g)] + torch.randn(3000, 2, generator=g) * 6
    wh = torch.rand(3000, 2, generator=g) * 80 + 8
    boxes = torch.cat((xy - wh / 2, xy + wh / 2), 1)
    scores = torch.rand(3000, generator=g)
    scores[100:120] = scores[100]  # ties keep their original order (stable sort)
    for thr in (0.3, 0.45, 0.7):
        want = nms_oracle.greedy_nms(boxes.numpy(), scores.numpy(), thr)
        got = nms_boxes(boxes.cuda(), scores.cuda(), thr).cpu().numpy()
        assert (want == got).all() and len(want) == len(got)
    assert nms_boxes(boxes[:0].cuda(), scores[:0].cuda(), 0.5).numel() == 0


@pytest.mark.parametrize("nms_type", ["batched_nms", "fast_nms", "matrix_nms", "merge_nms"])
@pytest.mark.parametrize("agnostic", [False, True])
def test_batched_nms_val2_other_types(nms_type, agnostic):
    """scripts/utils/nms.py:63-110: the val2 path's non-default nms_type branches on the batched kernels; selection identical
    to the oracle (pinned to the reference in tests/test_oracle_nms.py), decayed scores / merged boxes to fp32 rounding."""
    from ayolov2_b200.nms import batched_nms
    from oracle import nms_oracle

    pred = nms_oracle.synth_predictions(3, n=4000, nc=12, seed=13, cand_frac=0.2)
    pred[2, :, 4] = 0.0  # an image without candidates
    kw = dict(conf_thres=0.2, iou_thres=0.6, nms_box=400, agnostic=agnostic, nms_type=nms_type)
    want = nms_oracle.batched_nms(pred, **kw)
    got = batched_nms(pred.cuda(), **kw)
    assert sum(w.shape[0] for w in want) > 50 and want[2].shape[0] == 0
    for g, w in zip(got, want):
        assert g.shape == w.shape, (g.shape, w.shape)
        assert torch.equal(g[:, 5].cpu(), w[:, 5])
        if nms_type in ("batched_nms", "fast_nms"):
            assert torch.equal(g.cpu(), w)
        else:
            assert torch.allclose(g.cpu(), w, rtol=1e-5, atol=1e-4)


def test_batched_nms_val2_more_than_1024_survivors():
    """The reference has no bound on the survivors of the val2 path (nms.py:63-116); at val2's default conf 0.001 an image
    can keep more boxes than the batched kernel's 1024-entry kept list. Those images are finished by the unbounded route:
    nothing is truncated and nothing raises."""
    from ayolov2_b200.nms import batched_nms
    from oracle import nms_oracle

    g = torch.Generator().manual_seed(5)
    n, nc = 3000, 4
    pred = torch.zeros(2, n, 5 + nc)
    pred[..., :2] = torch.rand(2, n, 2, generator=g) * 2000      # spread out: almost nothing overlaps
    pred[..., 2:4] = 4.0 + 4.0 * torch.rand(2, n, 2, generator=g)
    pred[..., 4] = 0.5 + 0.5 * torch.rand(2, n, generator=g)
    pred[..., 5:] = torch.rand(2, n, nc, generator=g)
    pred[1, 200:, 4] = 0.0                                       # the second image stays below the bound
    for nms_type in ("nms", "merge_nms"):
        want = nms_oracle.batched_nms(pred, conf_thres=0.3, iou_thres=0.65, nms_box=2000, agnostic=True, nms_type=nms_type)
        got = batched_nms(pred.cuda(), conf_thres=0.3, iou_thres=0.65, nms_box=2000, agnostic=True, nms_type=nms_type)
        if nms_type == "nms":
            assert want[0].shape[0] > 1024 and want[1].shape[0] < 1024
        for a, b in zip(got, want):
            assert a.shape == b.shape
            assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=1e-4) if nms_type == "merge_nms" else torch.equal(a.cpu(), b)


def test_other_nms_types_more_than_max_nms_candidates():
    """metrics.py:378-379: above max_nms (30,000) candidates the rows are re-ranked by confidence and cut before the rule
    runs. Checks the re-ranked candidate table itself (rows, order, count, coordinate maximum; this also exercises the
    table-capacity retry, flags bit 0) and one whole rule ("batched_nms", whose class offset is that maximum + 1)."""
    import numpy as np

    from ayolov2_b200.nms import CandidateTable, non_max_suppression
    from oracle import nms_oracle

    g = torch.Generator().manual_seed(6)
    n, nc = 4200, 9
    pred = torch.zeros(1, n, 5 + nc)
    pred[..., :2] = torch.rand(1, n, 2, generator=g) * 600
    pred[..., 2:4] = 20.0 + 60.0 * torch.rand(1, n, 2, generator=g)
    pred[..., 4] = 0.6 + 0.4 * torch.rand(1, n, generator=g)
    pred[..., 5:] = 0.5 + 0.5 * torch.rand(1, n, nc, generator=g)   # every (row, class) pair passes: 37,800 candidates
    tab = CandidateTable(pred.cuda(), 0.25, True, None, max_nms=30000)
    assert tab.cap < n * nc and tab.settle() and int(tab.counts[0]) == 30000
    x = pred[0].numpy()
    conf = x[:, 5:] * x[:, 4:5]
    i, j = np.nonzero(conf > np.float32(0.25))
    rows = np.concatenate((nms_oracle.xywh2xyxy(x[i, :4]), conf[i, j, None], j[:, None].astype(np.float32)), 1)
    rows = rows[np.argsort(-rows[:, 4], kind="stable")[:30000]]
    assert np.array_equal(tab.rows[0, :30000, :6].cpu().numpy(), rows)
    assert float(tab.max_coord[0]) == float(rows[:, :4].max())
    kw = dict(conf_thres=0.25, iou_thres=0.5, multi_label=True, nms_type="batched_nms", max_det=300)
    want = nms_oracle.non_max_suppression(pred, **kw)
    got = non_max_suppression(pred.cuda(), **kw)
    assert want[0].shape[0] > 0 and torch.equal(got[0].cpu(), want[0])
