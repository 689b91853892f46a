"""CPU oracle for the input side of the path. TEST INFRASTRUCTURE ONLY (see oracle/nms_oracle.py for the import rule).

numpy restatement of
  * scripts/data_loader/data_loader.py:320-329   LoadImages._load_image after the decode (long side -> img_size)
  * scripts/data_loader/data_loader.py:395-459   LoadImages._letterbox (resize to the unpadded size, constant border)
  * scripts/data_loader/data_loader.py:388-389   HWC BGR -> CHW RGB (`img.transpose((2, 0, 1))[::-1]`)
  * scripts/data_loader/data_loader.py:461-477   LoadImages.collate_fn (torch.stack of the images)
  * scripts/data_loader/data_loader.py:888-909   LoadImagesAndLabels.collate_fn (image index into label column 0, cat)
`_letterbox` calls two functions of a third-party dependency that is not under /root/reference: OpenCV
(`opencv-python`, environment.yml; 4.13.0 is installed in the build container). Their published algorithm is restated
here in integer arithmetic:
  * cv2.resize(..., INTER_LINEAR) on 8-bit images (modules/imgproc/src/resize.cpp): per destination column
    fx = float((dx + 0.5) * scale_x - 0.5) with scale_x = 1 / (dst_w / src_w) in double, sx = floor(fx), fx -= sx, columns left of
    the image / at or past the last column collapse to one tap; the two taps are 11-bit fixed-point shorts
    round_half_even(w * 2048) (both rounded separately); rows likewise but clamped instead of collapsed; the
    horizontal pass keeps 32-bit integers, the vertical pass is
    (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    an exact 2x2 down-scale is rerouted to the area filter (a + b + c + d + 2) >> 2.
  * cv2.resize(..., INTER_AREA) on 8-bit images that shrink: see resize_area_u8.
  * cv2.copyMakeBorder(..., BORDER_CONSTANT, value=color).
Pinned bit-exact against cv2 itself on 400 random shape pairs, against the UNMODIFIED reference `_letterbox` /
`collate_fn` imported from /root/reference (tests/test_oracle_input.py, build container) and against the committed
fixture tests/golden/input_golden.npz (outputs of the unmodified reference, generator tests/golden/make_golden_input.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


def _linear_taps(ssize: int, dsize: int) -> Tuple[np.ndarray, np.ndarray]:
    """Source index and fractional weight of every destination index (resize.cpp: the `xofs` / `alpha` loop)."""
    inv_scale = np.float64(dsize) / np.float64(ssize)
    scale = np.float64(1.0) / inv_scale
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def _coef(w: np.ndarray) -> np.ndarray:
    """saturate_cast<short>(w * INTER_RESIZE_COEF_SCALE): round half to even."""
    return np.clip(np.rint((w.astype(np.float32) * np.float32(COEF_ONE)).astype(np.float32)), -32768, 32767).astype(np.int64)


def resize_linear_u8(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC images, bit-exact."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, _ = img.shape
    if w == 2 * dst_w and h == 2 * dst_h:  # INTER_LINEAR with an exact 2x2 decimation runs the area filter
        a = img.astype(np.int64)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, fx = _linear_taps(w, dst_w)
    left, right = sx < 0, sx >= w - 1
    fx[left], sx[left] = 0, 0
    fx[right], sx[right] = 0, w - 1
    a0, a1 = _coef(np.float32(1.0) - fx), _coef(fx)
    sx1 = np.minimum(sx + 1, w - 1)
    sy, fy = _linear_taps(h, dst_h)
    b0, b1 = _coef(np.float32(1.0) - fy), _coef(fy)
    y0, y1 = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
    src = img.astype(np.int64)
    rows = src[:, sx, :] * a0[None, :, None] + src[:, sx1, :] * a1[None, :, None]  # horizontal pass, [h, dst_w, c]
    r0, r1 = rows[y0], rows[y1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def _area_tab(ssize: int, dsize: int, scale: float):
    """resize.cpp computeResizeAreaTab: for every destination index the (source index, fp32 weight) entries, in order."""
    tabs = []
    for d in range(dsize):
        fs1 = np.float64(d) * scale
        fs2 = fs1 + scale
        cell = min(scale, np.float64(ssize) - fs1)
        s1, s2 = int(np.ceil(fs1)), int(np.floor(fs2))
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        e = []
        if s1 - fs1 > 1e-3:
            e.append((s1 - 1, np.float32((s1 - fs1) / cell)))
        for sidx in range(s1, s2):
            e.append((sidx, np.float32(1.0 / cell)))
        if fs2 - s2 > 1e-3:
            e.append((s2, np.float32(min(min(fs2 - s2, 1.0), cell) / cell)))
        tabs.append(e)
    width = max(len(e) for e in tabs)
    si = np.zeros((dsize, width), np.int64)
    al = np.zeros((dsize, width), np.float32)  # padding entries carry weight 0: x + 0 == x
    for d, e in enumerate(tabs):
        for t, (sidx, a) in enumerate(e):
            si[d, t], al[d, t] = sidx, a
    return si, al


def resize_area_u8(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_AREA) for uint8 HWC images that shrink in both directions,
    bit-exact. Integer ratios (resizeAreaFast_): integer cell sums, (a + b + c + d + 2) >> 2 for 2 x 2, otherwise
    round_half_even(sum * fp32(1 / area)); other ratios (resizeArea_): per source line the weighted fp32 sum of its cells in
    source order (buf += S * alpha), then the weighted fp32 sum of the lines (sum += beta * buf), rounded half to even."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, c = img.shape
    assert dst_w <= w and dst_h <= h
    f32 = np.float32
    scale_x, scale_y = 1.0 / (np.float64(dst_w) / w), 1.0 / (np.float64(dst_h) / h)
    isx, isy = int(np.rint(scale_x)), int(np.rint(scale_y))
    eps = np.finfo(np.float64).eps
    if abs(scale_x - isx) < eps and abs(scale_y - isy) < eps:
        a = img.astype(np.int64)
        if isx == 2 and isy == 2:
            return ((a[0:2 * dst_h:2, 0:2 * dst_w:2] + a[0:2 * dst_h:2, 1:2 * dst_w:2] + a[1:2 * dst_h:2, 0:2 * dst_w:2]
                     + a[1:2 * dst_h:2, 1:2 * dst_w:2] + 2) >> 2).astype(np.uint8)
        total = np.zeros((dst_h, dst_w, c), np.int64)
        for ky in range(isy):
            for kx in range(isx):
                total += a[ky:ky + dst_h * isy:isy, kx:kx + dst_w * isx:isx]
        v = (total.astype(f32) * (f32(1.0) / f32(isx * isy))).astype(f32)
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    xs, xa = _area_tab(w, dst_w, scale_x)
    ys, ya = _area_tab(h, dst_h, scale_y)
    src = img.astype(f32)
    buf = np.zeros((h, dst_w, c), f32)
    for t in range(xs.shape[1]):
        buf = (buf + (src[:, xs[:, t], :] * xa[None, :, t, None]).astype(f32)).astype(f32)
    out = np.zeros((dst_h, dst_w, c), f32)
    for t in range(ys.shape[1]):
        out = (out + (ya[:, t, None, None] * buf[ys[:, t]]).astype(f32)).astype(f32)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def load_image_resize(im: np.ndarray, img_size: int, augmentation: bool = False) -> np.ndarray:
    """data_loader.py:320-329 (`_load_image` after the decode): long side -> img_size; INTER_AREA when shrinking without
    augmentation, INTER_LINEAR otherwise; untouched when the long side already matches."""
    h0, w0 = im.shape[:2]
    r = img_size / max(h0, w0)
    if r == 1:
        return im
    dst_w, dst_h = int(w0 * r), int(h0 * r)
    if r < 1 and not augmentation:
        return resize_area_u8(im, dst_w, dst_h)
    return resize_linear_u8(im, dst_w, dst_h)


def letterbox_geometry(shape: Sequence[int], new_shape: Sequence[int], auto: bool = True, scale_fill: bool = False,
                       scale_up: bool = True, stride: int = 32):
    """data_loader.py:428-455 without the pixels: (new_unpad (w, h), ratio (w, h), (dw, dh), (top, bottom, left, right))."""
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scale_up:
        r = min(r, 1.0)
    ratio = (r, r)
    new_unpad = (int(round(shape[1] * r)), int(round(shape[0] * r)))
    dw = new_shape[1] - new_unpad[0]
    dh = new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    elif scale_fill:
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
        ratio = (new_shape[1] / shape[1], new_shape[0] / shape[0])
    dw = dw / 2
    dh = dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_unpad, ratio, (dw, dh), (top, bottom, left, right)


def letterbox(im: np.ndarray, new_shape: Sequence[int], color: Sequence[int] = (114, 114, 114), auto: bool = True,
              scale_fill: bool = False, scale_up: bool = True, stride: int = 32):
    """data_loader.py:395-459. Returns (image, ratio, (dw, dh)) like the reference."""
    shape = im.shape[:2]
    new_unpad, ratio, (dw, dh), (top, bottom, left, right) = letterbox_geometry(shape, new_shape, auto, scale_fill, scale_up, stride)
    if tuple(shape[::-1]) != tuple(new_unpad):
        im = resize_linear_u8(im, new_unpad[0], new_unpad[1])
    h, w = im.shape[:2]
    out = np.empty((h + top + bottom, w + left + right, im.shape[2]), dtype=np.uint8)
    out[...] = np.asarray(color, dtype=np.uint8)[None, None, :]
    out[top:top + h, left:left + w] = im
    return out, ratio, (dw, dh)


def to_chw_rgb(im: np.ndarray) -> np.ndarray:
    """data_loader.py:388-389: HWC BGR -> CHW RGB, contiguous."""
    return np.ascontiguousarray(im.transpose((2, 0, 1))[::-1])


def load_and_collate(images: List[np.ndarray], new_shape: Sequence[int], auto: bool = False, scale_fill: bool = False,
                     scale_up: bool = True, stride: int = 32, color: Sequence[int] = (114, 114, 114),
                     img_size: Optional[int] = None, augmentation: bool = False):
    """LoadImages.__getitem__ from the letterbox on (data_loader.py:380-393) for every image + collate_fn (:461-477):
    (uint8 [B, 3, H, W], shapes) with shapes[i] = ((h0, w0), ((h / h0, w / w0), (dw, dh))); images enter at their loaded
    size ((h0, w0) == (h, w)) or, with `img_size`, as decoded and go through `_load_image`'s resize first."""
    out, shapes = [], []
    for im in images:
        h0, w0 = im.shape[:2]
        if img_size is not None:  # decoded image: `_load_image`'s resize first (data_loader.py:320-329)
            im = load_image_resize(im, img_size, augmentation)
        lb, _, pad = letterbox(im, new_shape, color=color, auto=auto, scale_fill=scale_fill, scale_up=scale_up, stride=stride)
        out.append(to_chw_rgb(lb))
        h, w = im.shape[:2]
        shapes.append(((h0, w0), ((h / h0, w / w0), pad)))
    return np.stack(out, 0), tuple(shapes)


def collate_labels(labels: List[np.ndarray]) -> np.ndarray:
    """LoadImagesAndLabels.collate_fn (data_loader.py:905-909): column 0 = index of the image in the batch, rows concatenated."""
    out = []
    for i, l in enumerate(labels):
        l = np.array(l, dtype=np.float32, copy=True).reshape(-1, 6)
        l[:, 0] = i
        out.append(l)
    return np.concatenate(out, 0) if out else np.zeros((0, 6), np.float32)


def synth_images(seed: int, shapes: Sequence[Tuple[int, int]]) -> List[np.ndarray]:
    """Seeded BGR test images: smooth gradients + noise + a few rectangles (exercises the interpolation, not just noise)."""
    rng = np.random.default_rng(seed)
    out = []
    for (h, w) in shapes:
        yy, xx = np.mgrid[0:h, 0:w]
        base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 255 // max(h + w - 2, 1))], -1)
        im = (base * 0.6 + rng.integers(0, 103, (h, w, 3))).astype(np.uint8)
        for _ in range(3):
            y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
            im[y0:y0 + max(h // 5, 1), x0:x0 + max(w // 5, 1)] = rng.integers(0, 256, 3, dtype=np.uint8)
        out.append(np.ascontiguousarray(im))
    return out
