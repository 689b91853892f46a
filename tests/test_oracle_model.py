"""Pins for the kindle-compatible model layer and the CPU oracle forward: published parameter
counts, the 45/132 freeze split, fuse invariance, and (build container only) the fixture-weights -> detections
pin on the reference's own COCO val images."""
import glob
import os

import pytest
import torch

from ayolov2_b200 import synth
from oracle import nms_oracle, ref_import, yolo_oracle

REF = ref_import.REF_ROOT


@pytest.mark.parametrize("name,count", [("yolov5n", 1872157), ("yolov5s", 7235389), ("yolov5m", 21190557),
                                        ("yolov5l", 46563709), ("yolov5x", 86749405), ("yolov5_v5", 7276605)])
def test_param_counts(name, count):
    """README.md:206-211 of the reference (and 7,276,605 = unfused fixture graph)."""
    import kindle

    m = kindle.YOLOModel(synth.config_path(name))
    assert sum(p.numel() for p in m.parameters()) == count
    assert m.stride.tolist() == [8.0, 16.0, 32.0]
    assert m.output_save == [6, 4, 14, 10, 17, 20, 23]


def test_freeze_split_45_132():
    """tests/test_model_manager.py:60-61: freezing `model.{0..4}.` leaves 45 frozen / 132 trainable tensors."""
    import kindle

    m = kindle.YOLOModel(synth.config_path("yolov5s"))
    names = [n for n, _ in m.named_parameters()]
    frozen = [n for n in names if any(n.startswith(f"model.{i}.") for i in range(5))]
    assert len(frozen) == 45 and len(names) - len(frozen) == 132


def test_fuse_param_drop_and_invariance():
    """tests/test_tensor_decomposition.py:47 (7,276,605 -> 7,266,973) and tests/test_model_convert.py:43-44."""
    from copy import deepcopy

    m = synth.build_model("yolov5_v5", seed=0)
    x = torch.rand(1, 3, 128, 128)
    a = yolo_oracle.forward(m, x)[0]
    f = deepcopy(m).fuse()
    assert sum(p.numel() for p in f.parameters()) == 7266973
    b = yolo_oracle.forward(f, x)[0]
    assert torch.allclose(a, b, rtol=1e-3, atol=1e-4)


def test_eval_output_convention():
    m = synth.build_model("yolov5s", seed=0)
    pred, raw = yolo_oracle.forward(m, torch.rand(2, 3, 64, 96))
    assert pred.shape == (2, 3 * (8 * 12 + 4 * 6 + 2 * 3), 85)
    assert [tuple(r.shape) for r in raw] == [(2, 3, 8, 12, 85), (2, 3, 4, 6, 85), (2, 3, 2, 3, 85)]
    tr = yolo_oracle.forward(m, torch.rand(2, 3, 64, 96), training=True)
    assert isinstance(tr, list) and len(tr) == 3


def test_pickle_deepcopy_state_dict_names():
    import io
    from copy import deepcopy

    m = synth.build_model("yolov5s", seed=0)
    sd = m.state_dict()
    for k in ("model.0.conv.weight", "model.0.batch_norm.running_mean", "model.2.conv1.conv.weight",
              "model.2.bottleneck_c3.0.conv2.conv.weight", "model.24.conv.0.bias", "model.24.anchors", "model.24.anchor_grid"):
        assert k in sd, k
    buf = io.BytesIO()
    torch.save({"model": deepcopy(m).half()}, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)["model"].float()
    assert torch.equal(m2.state_dict()["model.24.anchors"], sd["model.24.anchors"].half().float())


def test_cpu_forward_is_refused():
    m = synth.build_model("yolov5n", seed=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 64, 64))


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "tests/res/weights/yolov5s_kindle.pt")), reason="reference fixture absent")
def test_fixture_checkpoint_detection_pin():
    """fixture weights through the restated operators + NMS on the 99 val images ->
    608 detections (+-1), 411 TP@0.5 (+-2) of 696 GT, mAP50 0.640. Also proves the reference's whole-module pickle loads through the kindle shim."""
    import cv2
    import kindle  # noqa: F401

    ck = torch.load(os.path.join(REF, "tests/res/weights/yolov5s_kindle.pt"), map_location="cpu", weights_only=False)
    m = ck["model"].float().eval()
    assert type(m).__name__ == "YOLOModel" and sum(p.numel() for p in m.parameters()) == 7276605
    import numpy as np

    from oracle import val_oracle

    ndet = ngt = ntp = 0
    stats = []
    for f in sorted(glob.glob(os.path.join(REF, "tests/res/datasets/coco/images/val2017/*.jpg"))):
        im = cv2.imread(f)
        h, w = im.shape[:2]
        r = 640 / max(h, w)
        im = cv2.resize(im, (round(w * r), round(h * r)), interpolation=cv2.INTER_LINEAR)
        nh, nw = im.shape[:2]
        top, left = (640 - nh) // 2, (640 - nw) // 2
        im = cv2.copyMakeBorder(im, top, 640 - nh - top, left, 640 - nw - left, cv2.BORDER_CONSTANT, value=(114, 114, 114))
        x = torch.from_numpy(im[:, :, ::-1].transpose(2, 0, 1).copy()).float()[None] / 255
        det = nms_oracle.non_max_suppression(yolo_oracle.forward(m, x)[0], 0.25, 0.45)[0].numpy()
        ndet += det.shape[0]
        # the reference-held ground truth of the same image (normalised cls cx cy w h) in the letterboxed 640 x 640 frame
        lf = f.replace("/images/", "/labels/").rsplit(".", 1)[0] + ".txt"
        lab = np.loadtxt(lf, ndmin=2).reshape(-1, 5) if os.path.isfile(lf) and os.path.getsize(lf) else np.zeros((0, 5))
        cx, cy, bw, bh = lab[:, 1] * nw + left, lab[:, 2] * nh + top, lab[:, 3] * nw, lab[:, 4] * nh
        lab_xyxy = np.stack((lab[:, 0], cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2), 1).astype(np.float32)
        correct = val_oracle.process_batch(det, lab_xyxy)  # train_utils.py:294-333 (class-aware, one detection per label)
        ngt += lab.shape[0]
        ntp += int(correct[:, 0].sum())
        stats.append((correct, det[:, 4], det[:, 5], lab[:, 0]))
    # 608 detections, 411 TP@IoU0.5 against 696 GT boxes, mAP50 0.640 / mAP50:95 0.475 -- the one
    # reference-held truth (its own weights, images and labels) that pins the restated operators' semantics
    assert abs(ndet - 608) <= 1, ndet
    assert ngt == 696, ngt
    assert abs(ntp - 411) <= 2, ntp
    tp, conf, pcls, tcls = (np.concatenate(a, 0) for a in zip(*stats))
    _, _, ap, _, _ = val_oracle.ap_per_class(tp, conf, pcls, tcls)
    assert abs(float(ap[:, 0].mean()) - 0.640) < 0.01 and abs(float(ap.mean()) - 0.475) < 0.01, (ap[:, 0].mean(), ap.mean())
