"""GPU input side (csrc/letterbox.cu through the C-ABI, ayolov2_b200/data_loader.py) against the oracle
(oracle/input_oracle.py, pinned to cv2 and to the unmodified reference's _letterbox / collate_fn) and against the committed
outputs of the unmodified reference (tests/golden/input_golden.npz). Integer work: the bar is bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu

from ayolov2_b200 import data_loader as dl  # noqa: E402
from ayolov2_b200 import ops  # noqa: E402
from oracle import input_oracle  # noqa: E402
from _parity import record  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "input_golden.npz")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_input import CASES  # noqa: E402


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_letterbox_collate_equals_reference_golden(ci):
    g = np.load(GOLD)
    new_shape, shapes, kw = CASES[ci]
    imgs = [g[f"c{ci}_img{k}"] for k in range(len(shapes))]
    kw = {k: v for k, v in kw.items() if k != "auto"}
    pb = dl.pack_batch(imgs, new_shape, **kw)
    out = pb.to_device("cuda")
    torch.cuda.synchronize()
    assert out.dtype == torch.uint8 and tuple(out.shape) == g[f"c{ci}_batch"].shape
    diff = np.abs(out.cpu().numpy().astype(np.int32) - g[f"c{ci}_batch"].astype(np.int32))
    record(f"input/letterbox_collate_vs_reference_golden_c{ci}", max_abs_diff=int(diff.max()), mismatching_bytes=int((diff > 0).sum()),
           bytes=int(diff.size))
    assert np.array_equal(out.cpu().numpy(), g[f"c{ci}_batch"])
    geo = g[f"c{ci}_geo"]
    for i in range(len(imgs)):
        assert pb.ratios[i] == (geo[i, 0], geo[i, 1]) and pb.shapes[i][1][1] == (geo[i, 2], geo[i, 3])


@pytest.mark.parametrize("new_shape", [(640, 640), (384, 672), (32, 36)])
def test_letterbox_collate_equals_oracle_ragged_shapes(new_shape):
    """Up- and down-scales, the exact 2 x 2 decimation, images already at the output size, one-pixel-wide / one-row images."""
    rng = np.random.default_rng(hash(new_shape) % 1000)
    H, W = new_shape
    shapes = [(H, W), (2 * H, 2 * W), (H, W // 2), (H // 2, W), (1, 7), (9, 1), (H - 1, W - 1), (H + 1, W + 3), (3 * H, 5 * W // 2)]
    shapes += [(int(rng.integers(2, 3 * H)), int(rng.integers(2, 3 * W))) for _ in range(7)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    for kw in (dict(), dict(scale_up=False), dict(scale_fill=True)):
        ref, ref_shapes = input_oracle.load_and_collate(imgs, new_shape, auto=False, **kw)
        pb = dl.pack_batch(imgs, new_shape, **kw)
        got = pb.to_device("cuda").cpu().numpy()
        bad = [i for i in range(len(imgs)) if not np.array_equal(got[i], ref[i])]
        record(f"input/letterbox_vs_oracle_{new_shape[0]}x{new_shape[1]}_{'_'.join(kw) or 'default'}", images=len(imgs),
               mismatching_bytes=int((got != ref).sum()), bytes=int(ref.size))
        assert not bad, f"images {[(i, shapes[i]) for i in bad]} differ ({kw})"
        assert pb.shapes == ref_shapes
        import dataclasses
        unknown = dataclasses.replace(pb, kinds=0)  # a C caller that does not classify its images: both kernels run
        assert np.array_equal(unknown.to_device("cuda").cpu().numpy(), ref)


def test_fused_space_to_depth_equals_two_step_path():
    """AY2_LB_S2D_BF16 == letterbox to uint8 NCHW followed by ay2_space_to_depth (prepare_img's /255 included), bit for bit,
    and the padding columns of the stem's input stay untouched."""
    H, W, B = 128, 160, 5
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (int(rng.integers(20, 300)), int(rng.integers(20, 300)), 3), dtype=np.uint8) for _ in range(B)]
    pb = dl.pack_batch(imgs, (H, W), pin=True)
    u8 = pb.to_device("cuda")
    Wp = W // 2 + 8
    ref = ops.ActView(torch.full((B, H // 2, Wp, 16), 7.0, dtype=torch.bfloat16, device="cuda"), 0, 16)
    got = ops.ActView(torch.full((B, H // 2, Wp, 16), 7.0, dtype=torch.bfloat16, device="cuda"), 0, 16)
    ops.space_to_depth(u8, ref, 1.0 / 255.0, x_offset=1)
    staging = torch.empty(pb.arena.numel() + 64, dtype=torch.uint8, device="cuda")
    pb.to_space_to_depth(got, 1.0 / 255.0, x_offset=1, staging=staging)
    torch.cuda.synchronize()
    assert torch.equal(got.buf.view(torch.int16), ref.buf.view(torch.int16))
    assert torch.all(got.buf[:, :, 0, :] == 7.0) and torch.all(got.buf[:, :, W // 2 + 1:, :] == 7.0)


def test_full_size_batch_properties():
    """BASELINE-size batch (64 x 3 x 640 x 640): size-independent properties instead of the (slow) oracle -- the border is
    the constant colour, an image already at the output size passes through unchanged (flipped to RGB planes), and the
    kernel is idempotent on its own output fed back as an image."""
    B, H, W = 64, 640, 640
    rng = np.random.default_rng(11)
    shapes = [(640, 640) if i % 4 == 0 else (640, int(rng.integers(300, 640))) if i % 4 == 1 else (int(rng.integers(300, 640)), 640)
              if i % 4 == 2 else (int(rng.integers(100, 400)), int(rng.integers(100, 400))) for i in range(B)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    pb = dl.pack_batch(imgs, (H, W), pin=True)
    out = pb.to_device("cuda")
    host = out.cpu().numpy()
    for i, im in enumerate(imgs):
        (uw, uh), _, _, (top, bottom, left, right) = dl.letterbox_geometry(im.shape[:2], (H, W), auto=False)
        mask = np.ones((H, W), bool)
        mask[top:top + uh, left:left + uw] = False
        assert np.all(host[i][:, mask] == 114)
        if im.shape[:2] == (uh, uw):
            assert np.array_equal(host[i][:, top:top + uh, left:left + uw], im.transpose(2, 0, 1)[::-1])
    again = dl.pack_batch([np.ascontiguousarray(host[i].transpose(1, 2, 0)[:, :, ::-1]) for i in range(B)], (H, W)).to_device("cuda")
    assert torch.equal(again, out)
    sample = [1, 2, 3, 62, 63]  # and five images against the oracle
    ref, _ = input_oracle.load_and_collate([imgs[i] for i in sample], (H, W))
    assert np.array_equal(host[sample], ref)


def test_collate_labels_equals_golden_and_oracle():
    g = np.load(GOLD)
    labels = [torch.from_numpy(g[f"lab_in{i}"]) for i in range(4)]
    got = dl.collate_labels(labels, "cuda").cpu().numpy()
    assert np.array_equal(got, g["lab_out"])
    rng = np.random.default_rng(5)
    many = [rng.random((int(n), 6)).astype(np.float32) for n in rng.integers(0, 40, 64)]
    got = dl.collate_labels([torch.from_numpy(m) for m in many], "cuda").cpu().numpy()
    assert np.array_equal(got, input_oracle.collate_labels(many))
    assert dl.collate_labels([torch.zeros(0, 6)] * 3, "cuda").shape == (0, 6)


def test_detector_from_loaded_images_equals_detector_on_the_reference_batch():
    """Detector.submit_packed (raw loaded images -> device letterbox, fused into the stem's input or through the uint8 slot)
    returns exactly the detections of Detector.detect on the batch the reference's Dataset + collate_fn would have built."""
    from ayolov2_b200 import synth
    from ayolov2_b200.detector import Detector

    model = synth.build_model("yolov5s", seed=2).cuda()
    B, H, W = 4, 320, 352
    imgs = input_oracle.synth_images(21, [(320, 352), (200, 352), (320, 240), (97, 131)])
    ref_batch, ref_shapes = input_oracle.load_and_collate(imgs, (H, W))
    host = torch.from_numpy(ref_batch)
    sample = host.cuda().float() / 255.0
    synth.calibrate_head(model, lambda: model(sample)[1], cand_frac=0.1)
    det = Detector(model, B, H, W, conf_thres=0.25, iou_thres=0.45, in_dtype=torch.uint8)
    want = det.detect(host.pin_memory())
    assert sum(x.shape[0] for x in want) > 20
    pb = dl.pack_batch(imgs, (H, W), pin=True)
    assert pb.shapes == ref_shapes
    for fused in (True, False, True):
        got = det.collect(det.submit_packed(pb, fused=fused))
        for a, b in zip(got, want):
            assert torch.equal(a, b)
    # pipelined: several batches in flight over the slots
    ks = [det.submit_packed(pb) for _ in range(2)]
    for k in ks:
        for a, b in zip(det.collect(k), want):
            assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
@pytest.mark.parametrize("shape,size", [((2, 3, 128, 128), (96, 96)), ((3, 3, 64, 96), (160, 224)), ((1, 3, 640, 640), (352, 352)),
                                        ((2, 3, 33, 47), (64, 32))])
def test_resize_bilinear_matches_torch_interpolate(dtype, shape, size):
    """ay2_resize_bilinear == F.interpolate(prepare_img(x), size, bilinear, align_corners=False) (yolo_trainer.py:223-248 after
    abstract_trainer.py:252-261). Floating point: same formula and operation order as torch's kernel; asserted to 2e-6 absolute
    on values in [0, 1] (a few fp32 ulps; north_star's fp32 tolerance is 1e-3), measured error recorded."""
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(sum(shape) + size[0])
    if dtype == torch.uint8:
        x = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8).cuda()
        ref = F.interpolate(x.float() / 255.0, size=size, mode="bilinear", align_corners=False)
        pre = 1.0 / 255.0
    else:
        x = torch.rand(shape, generator=g).cuda()
        ref = F.interpolate(x, size=size, mode="bilinear", align_corners=False)
        pre = 1.0
    out = torch.empty((shape[0], 3, *size), dtype=torch.float32, device="cuda")
    ops.resize_bilinear(x, out, pre)
    err = float((out - ref).abs().max())
    record(f"input/resize_bilinear_{'u8' if dtype == torch.uint8 else 'f32'}_{shape[2]}x{shape[3]}_to_{size[0]}x{size[1]}", max_abs=err,
           exact_fraction=float((out == ref).float().mean()))
    assert err <= 2e-6, err


def test_multi_scale_step_fills_the_engine_input_in_one_pass():
    """forward_train_resized(model, uint8 batch, size, 1/255) == model(F.interpolate(batch.float() / 255, size)) in train
    mode: same engine, same static input up to the resize's fp32 rounding, so the head outputs agree to bf16 noise."""
    import torch.nn.functional as F

    from ayolov2_b200 import synth
    from ayolov2_b200.train_engine import forward_train_resized

    model = synth.build_model("yolov5n", seed=0).cuda().train()
    x = torch.randint(0, 256, (4, 3, 128, 128), generator=torch.Generator().manual_seed(3), dtype=torch.uint8).cuda()
    size = (96, 160)
    want_in = F.interpolate(x.float() / 255.0, size=size, mode="bilinear", align_corners=False)
    with torch.no_grad():
        ref = [o.clone() for o in model(want_in)]
        eng = model.__dict__["_train_engine_last"]
        got = [o.clone() for o in forward_train_resized(model, x, size, 1.0 / 255.0)]
    assert model.__dict__["_train_engine_last"] is eng, "the resized route must reuse the engine of that shape"
    assert float((eng.static_in - want_in).abs().max()) <= 2e-6
    for a, b in zip(got, ref):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 2e-2 * float(b.abs().max())
    # and it trains: the gradient flows through the one autograd node
    outs = forward_train_resized(model, x, size, 1.0 / 255.0)
    sum((o * o).mean() for o in outs).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters() if p.requires_grad)


def test_candidate_overflow_redo_keeps_the_packed_input():
    """collect() redoes a batch whose candidate list overflowed (multi_label at a low conf_thres). For a batch submitted as
    loaded images the redo must start again from the raw-image arena of the slot (the uint8 slot holds nothing for it)."""
    from ayolov2_b200 import synth
    from ayolov2_b200.detector import Detector

    model = synth.build_model("yolov5s", seed=2).cuda()
    B, H, W = 2, 256, 256
    imgs = input_oracle.synth_images(31, [(256, 200), (120, 256)])
    ref_batch, _ = input_oracle.load_and_collate(imgs, (H, W))
    host = torch.from_numpy(ref_batch)
    sample = host.cuda().float() / 255.0
    synth.calibrate_head(model, lambda: model(sample)[1], cand_frac=0.1)
    kw = dict(conf_thres=0.05, iou_thres=0.45, multi_label=True, in_dtype=torch.uint8)
    want = Detector(model, B, H, W, **kw).detect(host.pin_memory())
    assert sum(x.shape[0] for x in want) > 20
    small = Detector(model, B, H, W, **kw)
    n, no = small.engine.pred.shape[1], small.engine.pred.shape[2]
    pb = dl.pack_batch(imgs, (H, W), pin=True)
    for fused in (True, False):
        small.nms_ws = ops.NmsWorkspace(B, n, no, max_det=300, multi_label=True, max_candidates=64, device=small.device)
        small._arm_candidates()
        small._graph, small._warm = None, False
        got = small.collect(small.submit_packed(pb, fused=fused))
        assert small.nms_ws.p.max_candidates > 64, "the overflow must have been detected and the workspace grown"
        for a, b in zip(got, want):
            assert torch.equal(a, b)


from make_golden_input import LOAD_CASES  # noqa: E402


def _resized_on_device(pb, staging):
    """The images the resize kernel wrote into the scratch space, in table order."""
    raw = pb.arena.numpy()
    lstart = pb.table_bytes + pb.load_table_offset
    lrec = raw[lstart:lstart + pb.n_load * dl._LOAD_REC.itemsize].view(dl._LOAD_REC)
    dev = staging.cpu().numpy()
    out = []
    for r in lrec:
        start = pb.table_bytes + int(r["dst_offset"])
        out.append(dev[start:start + 3 * int(r["dst_h"]) * int(r["dst_w"])].reshape(int(r["dst_h"]), int(r["dst_w"]), 3))
    return out


def test_load_resize_equals_reference_golden():
    """`_load_image`'s resize on the device (ay2_load_resize: INTER_AREA general / integer / 2 x 2, INTER_LINEAR) == the outputs
    of the unmodified reference (PNG -> LoadImages._load_image), and the collated batch == the oracle's composition."""
    g = np.load(GOLD)
    for li, ((h, w), size, aug) in enumerate(LOAD_CASES):
        im = g[f"load{li}_in"]
        pb = dl.pack_batch([im], (size, size), img_size=size, augmentation=aug)
        staging = torch.zeros(pb.host_bytes + pb.scratch_bytes, dtype=torch.uint8, device="cuda")
        out = pb.to_device(staging=staging)
        torch.cuda.synchronize()
        want = g[f"load{li}_out"]
        if pb.n_load:
            got = _resized_on_device(pb, staging)[0]
            diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
            record(f"input/load_resize_vs_reference_golden_{li}", max_abs_diff=int(diff.max()), mismatching_bytes=int((diff > 0).sum()),
                   bytes=int(diff.size))
            assert np.array_equal(got, want), (li, (h, w), size, aug)
        ref, ref_shapes = input_oracle.load_and_collate([im], (size, size), img_size=size, augmentation=aug)
        assert np.array_equal(out.cpu().numpy(), ref) and pb.shapes == ref_shapes


@pytest.mark.parametrize("img_size,aug", [(640, False), (96, False), (96, True)])
def test_decoded_batch_equals_oracle(img_size, aug):
    """Ragged decoded images -> collated batch: strong and weak shrinks, integer ratios (2, 3, 4), exact halves, up-scales,
    images already at the size; against the oracle (pinned to cv2 and the reference)."""
    rng = np.random.default_rng(img_size + aug)
    S = img_size
    shapes = [(S, S), (2 * S, 2 * S), (3 * S, 3 * S // 2), (4 * S, 4 * S), (2 * S, S), (S // 2, S // 3), (S + 1, S - 7), (5 * S // 2, 7 * S // 3),
              (S, S // 2), (S - 1, S - 1)]
    shapes += [(int(rng.integers(8, 3 * S)), int(rng.integers(8, 3 * S))) for _ in range(6)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    ref, ref_shapes = input_oracle.load_and_collate(imgs, (S, S), img_size=S, augmentation=aug)
    pb = dl.pack_batch(imgs, (S, S), img_size=S, augmentation=aug, pin=True)
    got = pb.to_device("cuda").cpu().numpy()
    bad = [(i, shapes[i]) for i in range(len(imgs)) if not np.array_equal(got[i], ref[i])]
    record(f"input/decoded_batch_vs_oracle_{S}_{'aug' if aug else 'val'}", images=len(imgs), mismatching_bytes=int((got != ref).sum()),
           bytes=int(ref.size))
    assert not bad, bad
    assert pb.shapes == ref_shapes


def test_detector_from_decoded_images():
    from ayolov2_b200 import synth
    from ayolov2_b200.detector import Detector

    model = synth.build_model("yolov5s", seed=2).cuda()
    B, S = 3, 256
    imgs = input_oracle.synth_images(41, [(512, 384), (300, 411), (256, 256)])
    ref_batch, ref_shapes = input_oracle.load_and_collate(imgs, (S, S), img_size=S)
    host = torch.from_numpy(ref_batch)
    sample = host.cuda().float() / 255.0
    synth.calibrate_head(model, lambda: model(sample)[1], cand_frac=0.1)
    det = Detector(model, B, S, S, conf_thres=0.25, iou_thres=0.45, in_dtype=torch.uint8)
    want = det.detect(host.pin_memory())
    pb = dl.pack_batch(imgs, (S, S), img_size=S, pin=True)
    assert pb.shapes == ref_shapes and pb.n_load == 2
    for fused in (True, False):
        for a, b in zip(det.collect(det.submit_packed(pb, fused=fused)), want):
            assert torch.equal(a, b)
