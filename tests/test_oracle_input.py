"""Pins oracle/input_oracle.py (letterbox + collate, data_loader.py:388-477,888-909): against the committed outputs of the
unmodified reference (tests/golden/input_golden.npz), against the reference itself when /root/reference is present, and
the restated cv2.resize against cv2 when it is importable. CPU only."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import input_oracle, ref_import  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "input_golden.npz")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_input import CASES  # noqa: E402

HAVE_CV2 = importlib.util.find_spec("cv2") is not None


def golden_case(g, ci):
    new_shape, shapes, kw = CASES[ci]
    imgs = [g[f"c{ci}_img{k}"] for k in range(len(shapes))]
    return new_shape, imgs, kw, g[f"c{ci}_batch"], g[f"c{ci}_geo"]


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_letterbox_collate_matches_golden(ci):
    g = np.load(GOLD)
    new_shape, imgs, kw, batch, geo = golden_case(g, ci)
    for im, (h, w) in zip(imgs, CASES[ci][1]):
        assert im.shape == (h, w, 3)
    got, shapes = input_oracle.load_and_collate(imgs, new_shape, **kw)
    assert got.dtype == np.uint8 and got.shape == batch.shape
    assert np.array_equal(got, batch)
    for im, row, shp in zip(imgs, geo, shapes):
        _, ratio, pad, _ = input_oracle.letterbox_geometry(im.shape[:2], new_shape, **kw)
        assert tuple(row) == (ratio[0], ratio[1], float(pad[0]), float(pad[1]))
        assert shp[0] == im.shape[:2] and shp[1][1] == pad


def test_synth_images_reproduce_the_fixture_inputs():
    g = np.load(GOLD)
    for ci, (_, shapes, _) in enumerate(CASES):
        for k, im in enumerate(input_oracle.synth_images(100 + ci, shapes)):
            assert np.array_equal(im, g[f"c{ci}_img{k}"])


def test_collate_labels_matches_golden():
    g = np.load(GOLD)
    labels = [g[f"lab_in{i}"] for i in range(4)]
    got = input_oracle.collate_labels(labels)
    assert np.array_equal(got, g["lab_out"])
    assert got[:, 0].tolist() == [0, 0, 0, 2, 2, 2, 2, 2, 3]
    assert input_oracle.collate_labels([]).shape == (0, 6)


def test_auto_mode_pads_to_the_stride_multiple():
    im = input_oracle.synth_images(3, [(100, 60)])[0]
    out, ratio, (dw, dh) = input_oracle.letterbox(im, (128, 128), auto=True, stride=32)
    assert out.shape[0] == 128 and out.shape[1] % 32 == 0 and out.shape[1] < 128
    assert np.all(out[:, 0] == 114) or dw < 1


@pytest.mark.skipif(not HAVE_CV2, reason="cv2 not installed")
def test_resize_restatement_is_bit_exact_against_cv2():
    import cv2

    rng = np.random.default_rng(0)
    for t in range(150):
        h, w = (int(v) for v in rng.integers(2, 300, 2))
        if t % 7 == 0:
            dh, dw = max(h // 2, 1), max(w // 2, 1)
            h, w = 2 * dh, 2 * dw
        elif t % 7 == 1:
            dh, dw = h, int(rng.integers(1, 300))
        else:
            dh, dw = (int(v) for v in rng.integers(1, 300, 2))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref, input_oracle.resize_linear_u8(img, dw, dh)), ((h, w), (dh, dw))


@pytest.mark.skipif(not (ref_import.available() and HAVE_CV2), reason="reference tree / cv2 not present")
def test_oracle_matches_the_unmodified_reference():
    dl = ref_import.load_data_loader()
    fake = types.SimpleNamespace(img_size=160, stride=32)
    rng = np.random.default_rng(1)
    for case in range(12):
        shapes = [(int(rng.integers(20, 260)), int(rng.integers(20, 260))) for _ in range(3)]
        kw = [dict(auto=False), dict(auto=True), dict(auto=False, scale_up=False), dict(auto=False, scale_fill=True)][case % 4]
        for im in input_oracle.synth_images(case, shapes):
            ref, r_ratio, r_pad = dl.LoadImages._letterbox(fake, im, new_shape=(160, 128), **kw)
            got, ratio, pad = input_oracle.letterbox(im, (160, 128), **kw)
            assert np.array_equal(ref, got), (im.shape, kw)
            assert tuple(r_ratio) == tuple(ratio) and tuple(r_pad) == tuple(pad)


from make_golden_input import LOAD_CASES  # noqa: E402


@pytest.mark.parametrize("li", range(len(LOAD_CASES)))
def test_load_image_resize_matches_reference_golden(li):
    """`_load_image` after the decode (data_loader.py:320-329): outputs of the unmodified reference on PNG files."""
    g = np.load(GOLD)
    (h, w), size, aug = LOAD_CASES[li]
    im = g[f"load{li}_in"]
    assert im.shape == (h, w, 3)
    got = input_oracle.load_image_resize(im, size, aug)
    assert got.shape == g[f"load{li}_out"].shape and np.array_equal(got, g[f"load{li}_out"])


@pytest.mark.skipif(not HAVE_CV2, reason="cv2 not installed")
def test_area_restatement_is_bit_exact_against_cv2():
    import cv2

    rng = np.random.default_rng(2)
    for t in range(120):
        h, w = (int(v) for v in rng.integers(8, 260, 2))
        if t % 5 == 0:  # integer ratios, equal and unequal
            ky, kx = [(2, 2), (3, 3), (2, 3), (4, 2), (1, 2), (5, 5)][(t // 5) % 6]
            dh, dw = max(h // ky, 1), max(w // kx, 1)
            h, w = dh * ky, dw * kx
        elif t % 5 == 1:  # what _load_image asks for
            r = 64 / max(h, w)
            if r >= 1:
                continue
            dw, dh = int(w * r), int(h * r)
        else:
            dh, dw = int(rng.integers(1, h + 1)), int(rng.integers(1, w + 1))
        if (dh, dw) == (h, w) or dh < 1 or dw < 1:
            continue
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA)
        assert np.array_equal(ref, input_oracle.resize_area_u8(img, dw, dh)), ((h, w), (dh, dw))


@pytest.mark.skipif(not (ref_import.available() and HAVE_CV2), reason="reference tree / cv2 not present")
def test_load_image_oracle_matches_the_unmodified_reference(tmp_path):
    from make_golden_input import reference_load_image

    dl = ref_import.load_data_loader()
    rng = np.random.default_rng(3)
    for case in range(10):
        h, w = (int(v) for v in rng.integers(20, 300, 2))
        size = int(rng.choice([64, 96, 160]))
        aug = case % 3 == 2
        im = input_oracle.synth_images(400 + case, [(h, w)])[0]
        ref, _, _ = reference_load_image(dl, im, size, aug, str(tmp_path))
        got = input_oracle.load_image_resize(im, size, aug)
        assert ref.shape == got.shape and np.array_equal(ref, got), ((h, w), size, aug)


def _area_cells(d, ssize, scale):
    """Per-pixel form of computeResizeAreaTab, as csrc/letterbox.cu (area_cells) evaluates it on the device."""
    import math

    fs1 = np.float64(d) * np.float64(scale)
    fs2 = fs1 + np.float64(scale)
    cell = min(np.float64(scale), np.float64(ssize) - fs1)
    s1, s2 = int(math.ceil(fs1)), int(math.floor(fs2))
    s2 = min(s2, ssize - 1)
    s1 = min(s1, s2)
    dl, dr = np.float64(s1) - fs1, fs2 - np.float64(s2)
    f32 = np.float32
    return s1, s2, dl > 1e-3, dr > 1e-3, f32(dl / cell), f32(np.float64(1.0) / cell), f32(min(min(dr, 1.0), cell) / cell)


def _area_pixel(img, dx, dy, sx, sy):
    """One output pixel of load_resize_kernel's general INTER_AREA branch: lines in order, cells in order, separate fp32
    multiplies and adds."""
    f32 = np.float32
    h, w, _ = img.shape
    x1, x2, xl, xr, xwl, xwm, xwr = _area_cells(dx, w, sx)
    y1, y2, yl, yr, ywl, ywm, ywr = _area_cells(dy, h, sy)
    cols = ([(x1 - 1, xwl)] if xl else []) + [(c, xwm) for c in range(x1, x2)] + ([(x2, xwr)] if xr else [])
    rows = ([(y1 - 1, ywl)] if yl else []) + [(r, ywm) for r in range(y1, y2)] + ([(y2, ywr)] if yr else [])
    total = np.zeros(3, f32)
    for r, beta in rows:
        buf = np.zeros(3, f32)
        for c, alpha in cols:
            buf = (buf + (img[r, c].astype(f32) * alpha).astype(f32)).astype(f32)
        total = (total + (beta * buf).astype(f32)).astype(f32)
    return np.clip(np.rint(total), 0, 255).astype(np.uint8)


def test_per_pixel_area_formulation_equals_the_table_formulation():
    """The device computes every output pixel on its own (cells and weights from the pixel index) while OpenCV and the oracle
    build tables first: both must give the same bytes (this pins the kernel's algorithm on the CPU; the kernel itself is
    compared with the oracle in tests/test_input_gpu.py)."""
    rng = np.random.default_rng(0)
    cases = [((30, 23), (13, 10)), ((31, 47), (30, 20)), ((17, 9), (17, 4)), ((50, 33), (7, 5)), ((40, 40), (39, 39)), ((9, 40), (3, 13))]
    cases += [((int(h), int(w)), (int(rng.integers(1, h + 1)), int(rng.integers(1, w + 1)))) for h, w in rng.integers(4, 40, (8, 2))]
    for (h, w), (dh, dw) in cases:
        sx, sy = 1.0 / (float(dw) / float(w)), 1.0 / (float(dh) / float(h))
        if (dh, dw) == (h, w) or (abs(sx - round(sx)) < 2.3e-16 and abs(sy - round(sy)) < 2.3e-16):
            continue  # integer ratios take the cell-sum path
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = input_oracle.resize_area_u8(img, dw, dh)
        got = np.stack([np.stack([_area_pixel(img, dx, dy, sx, sy) for dx in range(dw)], 0) for dy in range(dh)], 0)
        assert np.array_equal(got, want), ((h, w), (dh, dw))
