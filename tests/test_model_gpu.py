"""End-to-end parity of the compiled CUDA forward (ayolov2_b200.engine) with the CPU fp32 oracle
(oracle/yolo_oracle.py) on seeded random weights shared through the same module tree.

Tolerance (BASELINE.json north_star): bf16 activations -> 1e-2 relative, asserted BOTH as max-norm
(max|d| / max|ref|) and as relative L2 on the head logits and the decoded boxes; class / objectness probabilities
additionally within 1e-2 absolute. Every measured error is recorded (tests/_parity.py -> profiles/r02_parity.json;
round-2 measurements: logits 2.4e-3 .. 3.2e-3 max-norm, 1.8e-3 .. 1.9e-3 rel-L2). The fp32 clause of the north star
(1e-3) is tested through the split-bf16 engine mode in tests/test_precise_gpu.py."""
import pytest
import torch

from _parity import errs, record

pytestmark = pytest.mark.gpu
BF16_TOL = 1e-2  # BASELINE.json north_star: "bf16 within 1e-2"


def _norm_err(got, ref):
    return float((got.float().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-6))


def _assert_forward_parity(case, got_pred, got_raw, want_pred, want_raw, tol=BF16_TOL):
    assert got_pred.shape == want_pred.shape
    for i, (gr, wr) in enumerate(zip(got_raw, want_raw)):
        assert gr.shape == wr.shape
        e = errs(gr, wr)
        record(f"{case}/logits_P{i + 3}", **e)
        assert e["max_norm"] < tol and e["rel_l2"] < tol, (case, i, e)
    gp = got_pred.float().cpu()
    eb = errs(gp[..., :4], want_pred[..., :4])
    pabs = float((gp[..., 4:] - want_pred[..., 4:]).abs().max())
    record(f"{case}/decoded", box_max_norm=eb["max_norm"], box_rel_l2=eb["rel_l2"], prob_max_abs=pabs)
    assert eb["max_norm"] < tol and eb["rel_l2"] < tol, (case, eb)
    assert pabs < tol, (case, pabs)


@pytest.mark.parametrize("name,hw,B", [("yolov5s", (320, 320), 2), ("yolov5s", (256, 384), 2), ("yolov5_v5", (256, 256), 2),
                                       ("yolov5n", (192, 192), 2), ("yolov5s", (640, 640), 1), ("yolov5m", (320, 320), 1)])
def test_forward_matches_oracle(name, hw, B):
    """Includes BASELINE.json configs[0]: yolov5s.yaml forward on a 1 x 3 x 640 x 640 random tensor."""
    from ayolov2_b200 import synth as model_utils
    from oracle import yolo_oracle

    model = model_utils.build_model(name, seed=0)
    g = torch.Generator().manual_seed(5)
    x = torch.rand((B, 3, *hw), generator=g)
    want_pred, want_raw = yolo_oracle.forward(model, x)
    model_cuda = model.cuda()
    got_pred, got_raw = model_cuda(x.cuda())
    torch.cuda.synchronize()
    _assert_forward_parity(f"forward/{name}_{hw[0]}x{hw[1]}_b{B}", got_pred, got_raw, want_pred, want_raw)


def test_uint8_detector_matches_float_path():
    """Detector (uint8 input, /255 inside the space-to-depth kernel, CUDA graph, NMS) == Engine on float input."""
    from ayolov2_b200.detector import Detector
    from ayolov2_b200.nms import non_max_suppression
    from ayolov2_b200 import synth as model_utils

    model = model_utils.build_model("yolov5s", seed=1).cuda()
    B, H, W = 2, 320, 320
    g = torch.Generator().manual_seed(2)
    img8 = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8)
    det = Detector(model, B, H, W, conf_thres=0.001, iou_thres=0.6, in_dtype=torch.uint8)
    out1 = det.detect(img8.pin_memory())
    out2 = det.detect(img8.pin_memory())  # graph replay path
    pred, _ = model(img8.cuda().float() / 255.0)
    want = non_max_suppression(pred, 0.001, 0.6)
    for a, b, c in zip(out1, out2, want):
        assert torch.equal(a, b)
        assert a.shape == c.shape and torch.allclose(a, c.cpu(), atol=1e-3, rtol=1e-3)


def test_fuse_invariance():
    """tests/test_model_convert.py:43-44 of the reference: model(x)[0] == model.fuse()(x)[0]."""
    from copy import deepcopy

    from ayolov2_b200 import synth as model_utils

    model = model_utils.build_model("yolov5s", seed=3).cuda()
    x = torch.rand((1, 3, 256, 256), device="cuda")
    a = model(x)[0]
    fused = deepcopy(model).fuse()
    n0 = sum(p.numel() for p in model.parameters())
    n1 = sum(p.numel() for p in fused.parameters())
    assert n0 - n1 == sum(m.num_features for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d))
    b = fused(x)[0]
    assert torch.allclose(a, b, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("multi_label", [False, True])
@pytest.mark.parametrize("hw", [(320, 320), (352, 288)])
def test_fused_head_nms_equals_dense_path(multi_label, hw):
    """Detector's fused head must select exactly what ay2_head_decode + ay2_nms_batched select (same arithmetic, no
    dense (B, 25200, 85) tensor), both with the candidates scored by the detect convolutions' epilogues
    (ay2_conv_plan_set_head_candidates + ay2_nms_from_candidates) and by the stand-alone kernels
    (ay2_nms_from_logits). 352x288 gives 11x9 / 22x18 / 44x36 maps: output tiles overhang the image."""
    from ayolov2_b200 import synth
    from ayolov2_b200.detector import Detector

    model = synth.build_model("yolov5s", seed=2).cuda()
    B, (H, W) = 3, hw
    img = torch.randint(0, 256, (B, 3, H, W), generator=torch.Generator().manual_seed(9), dtype=torch.uint8)
    sample = img.cuda().float() / 255.0
    synth.calibrate_head(model, lambda: model(sample)[1], cand_frac=0.1)
    kw = dict(conf_thres=0.25, iou_thres=0.45, multi_label=multi_label, in_dtype=torch.uint8)
    fused = Detector(model, B, H, W, **kw)
    logits = Detector(model, B, H, W, fuse_candidates=False, **kw)
    dense = Detector(model, B, H, W, dense_pred=True, **kw)
    assert fused.fused_candidates and not logits.fused_candidates
    a = fused.detect(img.pin_memory())
    a2 = fused.detect(img.pin_memory())  # graph replay: the counters are re-zeroed inside the graph
    l = logits.detect(img.pin_memory())
    b = dense.detect(img.pin_memory())
    assert sum(x.shape[0] for x in a) > 50, "calibration should produce detections"
    for x, x2, y, z in zip(a, a2, b, l):
        assert x.shape == y.shape and torch.equal(x, y)
        assert torch.equal(x, x2) and torch.equal(z, y)
    # the candidate sets themselves (unordered) are identical, not just the survivors
    nb = fused.nms_ws.p.batch
    ca = fused.nms_ws.ws[: 4 * nb].view(torch.int32).cpu()
    cl = logits.nms_ws.ws[: 4 * nb].view(torch.int32).cpu()
    assert torch.equal(ca, cl) and int(ca.sum()) > 0


def test_tucker_decomposed_forward_matches_oracle():
    """BASELINE.json configs[3]: Tucker-2 decomposed yolov5s (1x1 -> kxk -> 1x1 chains, decomposition.py:363-424) on the
    CUDA engine vs the nn.Sequential oracle; ranks here are odd sizes to exercise the zero-padding to 16."""
    from ayolov2_b200 import synth, tucker
    from oracle import yolo_oracle

    model = synth.build_model("yolov5s", seed=4)
    names = tucker.decompose_model_fixed(model, ratio=0.45)
    assert len(names) >= 15
    x = torch.rand((2, 3, 256, 256), generator=torch.Generator().manual_seed(1))
    want_pred, want_raw = yolo_oracle.forward(model, x)
    got_pred, got_raw = model.cuda()(x.cuda())
    torch.cuda.synchronize()
    # bf16 storage of the rank-R intermediates: same 1e-2 bound as the dense model (the 1e-3 clause of the north star
    # for the decomposed model is tested in fp32-equivalent arithmetic, tests/test_precise_gpu.py)
    _assert_forward_parity("tucker_bf16/yolov5s_256x256_r0.45", got_pred, got_raw, want_pred, want_raw)


@pytest.mark.parametrize("name,decomposed", [("yolov5s", False), ("yolov5s", True), ("yolov5m", False)])
def test_fused_chains_match_separate_launches(name, decomposed):
    """The fused chain kernel (Bottleneck = 1 launch, Tucker chain = 1 launch, C3 conv3 reading two sources) against the
    same engine with every link as its own launch: identical bf16 rounding points, so the logits agree to accumulation
    order. 640x640 so that the 160x160 / 80x80 / 40x40 maps (the fused ones) are all exercised."""
    from ayolov2_b200 import engine as eng_mod, synth, tucker

    model = synth.build_model(name, seed=3)
    if decomposed:
        tucker.decompose_model_fixed(model, ratio=0.5)
    model = model.cuda().eval()
    x = torch.rand((2, 3, 640, 640), generator=torch.Generator().manual_seed(9)).cuda()

    def run(fuse):
        eng_mod.Builder.FUSE_CHAINS = fuse
        keep = eng_mod.Builder.FUSE_BOTTLENECK_MAX_C
        eng_mod.Builder.FUSE_BOTTLENECK_MAX_C = 128  # exercise every fusable Bottleneck, not only the ones the policy picks
        try:
            e = eng_mod.Engine(model, 2, 640, 640, use_graph=False)
            pred, raw = e.run(x)
            torch.cuda.synchronize()
            nchain = sum(1 for p in e.b.plans if type(p).__name__ == "ChainPlan")
            return pred.clone(), [r.clone() for r in raw], nchain, len(e.b.steps)
        finally:
            eng_mod.Builder.FUSE_CHAINS = True
            eng_mod.Builder.FUSE_BOTTLENECK_MAX_C = keep

    p1, r1, n1, s1 = run(True)
    p0, r0, n0, s0 = run(False)
    assert n0 == 0 and n1 >= 8, (n0, n1)
    assert s1 < s0
    for i, (a, b) in enumerate(zip(r1, r0)):
        e = errs(a, b)
        record(f"fused_vs_separate/{name}{'_tucker' if decomposed else ''}/logits_P{i + 3}", **e)
        assert e["max_norm"] < 1e-2 and e["rel_l2"] < 1e-2, e


def test_tta_matches_oracle_views():
    """inference_with_tta (tta_utils.py:62-86; the reference's val.py scales 1 / 0.83 / 0.67 with a left-right flip on the
    second view): every augmented view runs through the CUDA engine; the concatenated, de-augmented prediction equals the
    same orchestration over the fp32 CPU oracle forward within the bf16 tolerance."""
    from ayolov2_b200 import synth as model_utils, tta
    from oracle import yolo_oracle

    model = model_utils.build_model("yolov5s", seed=3)
    x = torch.rand((2, 3, 256, 320), generator=torch.Generator().manual_seed(8))
    s, f = [1, 0.83, 0.67], [None, 3, None]

    class _Oracle(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m, self.model, self.stride = m, m.model, m.stride

        def forward(self, xi):
            return yolo_oracle.forward(self.m, xi)

    want, _ = tta.inference_with_tta(_Oracle(model), x, s, f)
    got, _ = tta.inference_with_tta(model.cuda(), x.cuda(), s, f)
    got = got.float().cpu()
    assert got.shape == want.shape
    eb = errs(got[..., :4], want[..., :4])
    pabs = float((got[..., 4:] - want[..., 4:]).abs().max())
    record("tta/yolov5s_256x320", box_max_norm=eb["max_norm"], box_rel_l2=eb["rel_l2"], prob_max_abs=pabs)
    assert eb["max_norm"] < BF16_TOL and eb["rel_l2"] < BF16_TOL and pabs < BF16_TOL
