"""Input side of the path on the GPU: letterbox + BGR->RGB + HWC->CHW + collate of a whole batch.

The reference does this per image on the host, inside the Dataset (scripts/data_loader/data_loader.py):
  `LoadImages.__getitem__` (:373-393)  -> `_letterbox` (:395-459, cv2.resize + cv2.copyMakeBorder) -> transpose + channel
  flip (:388-389), then `collate_fn` (:461-477 / :888-909) stacks the images and numbers the label rows.
Here the Dataset hands over the LOADED images as they are (ragged HWC BGR uint8); the work is split in two halves:

  * `pack_batch(images, new_shape, ...)` -- host, no pixel is touched except for one memcpy per image into a single
    byte arena (pinned on request). It computes the per-image geometry (`letterbox_geometry`: the scalar arithmetic of
    :428-455) and the reference's `shapes` tuples. Usable as a DataLoader `collate_fn` (pure CPU, picklable result).
  * `PackedBatch.to_device(...)` -- ONE host->device copy of the arena (the raw pixels are fewer bytes than the padded
    batch) and one call (`ay2_letterbox_collate`, csrc/letterbox.cu: a launch for the images that
    enter at their final size, one for those that are resized) that resizes (cv2's 8-bit INTER_LINEAR, bit-exact),
    pads, flips the channels and writes either the reference's collated uint8 NCHW tensor or directly the bf16
    space-to-depth image the stem convolution reads (then `prepare_img`'s /255 and `ay2_space_to_depth` are fused in).

`collate_labels` is `LoadImagesAndLabels.collate_fn`'s label half (:905-909) on the device.
There is no CPU fallback: `to_device` needs the CUDA library and raises without it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

LB_NCHW_U8, LB_S2D_BF16 = 0, 1
LB_HAS_COPY, LB_HAS_RESIZE = 1, 2
_REC = np.dtype([("src_offset", "<i8"), ("scale_x", "<f8"), ("scale_y", "<f8"), ("src_h", "<i4"), ("src_w", "<i4"), ("src_row_bytes", "<i4"), ("dst_h", "<i4"),
                 ("dst_w", "<i4"), ("top", "<i4"), ("left", "<i4"), ("reserved", "<i4")])  # ay2_letterbox_image
assert _REC.itemsize == 56
LR_LINEAR, LR_AREA = 0, 1
_LOAD_REC = np.dtype([("src_offset", "<i8"), ("dst_offset", "<i8"), ("scale_x", "<f8"), ("scale_y", "<f8"), ("src_h", "<i4"),
                      ("src_w", "<i4"), ("src_row_bytes", "<i4"), ("dst_h", "<i4"), ("dst_w", "<i4"), ("dst_row_bytes", "<i4"),
                      ("mode", "<i4"), ("reserved", "<i4")])  # ay2_load_resize_image
assert _LOAD_REC.itemsize == 64


def letterbox_geometry(shape: Sequence[int], new_shape: Sequence[int], auto: bool = True, scale_fill: bool = False,
                       scale_up: bool = True, stride: int = 32):
    """The geometry of `_letterbox` (data_loader.py:428-455) for a loaded image of `shape` (h, w): returns
    (new_unpad (w, h), ratio (w, h), (dw, dh), (top, bottom, left, right)). ratio and (dw, dh) are the values the
    reference returns (and `scale_coords` later consumes)."""
    h, w = int(shape[0]), int(shape[1])
    gain = min(new_shape[0] / h, new_shape[1] / w)
    if not scale_up:
        gain = min(gain, 1.0)
    ratio = (gain, gain)
    unpad_w, unpad_h = int(round(w * gain)), int(round(h * gain))
    pad_w, pad_h = new_shape[1] - unpad_w, new_shape[0] - unpad_h
    if auto:  # minimum rectangle: only the remainder modulo the stride is padded
        pad_w, pad_h = np.mod(pad_w, stride), np.mod(pad_h, stride)
    elif scale_fill:  # stretch to the full shape, no border
        pad_w, pad_h = 0.0, 0.0
        unpad_w, unpad_h = int(new_shape[1]), int(new_shape[0])
        ratio = (new_shape[1] / w, new_shape[0] / h)
    pad_w, pad_h = pad_w / 2, pad_h / 2
    border = (int(round(pad_h - 0.1)), int(round(pad_h + 0.1)), int(round(pad_w - 0.1)), int(round(pad_w + 0.1)))
    return (unpad_w, unpad_h), ratio, (pad_w, pad_h), border


@dataclass
class PackedBatch:
    """Host half of a batch: [table | images] in one uint8 tensor + what the reference's collate_fn returns beside the images."""

    arena: torch.Tensor          # uint8 [bytes]: `batch` ay2_letterbox_image records, then the images (16-byte aligned)
    batch: int
    out_shape: Tuple[int, int]   # (H, W) of the collated tensor
    shapes: tuple                # per image ((h0, w0), ((h / h0, w / w0), (dw, dh)))  -- data_loader.py:391
    ratios: tuple                # per image (ratio_w, ratio_h) returned by _letterbox
    color: Tuple[int, int, int] = (114, 114, 114)
    paths: tuple = ()
    kinds: int = 0               # LB_HAS_COPY | LB_HAS_RESIZE: which kernels the batch needs
    # decode-side resize (`_load_image`, data_loader.py:320-329) of `n_load` images: its table sits inside the arena, the
    # resized images are written to `scratch_bytes` of DEVICE memory right behind the uploaded bytes
    host_bytes: int = 0
    n_load: int = 0
    load_table_offset: int = 0   # relative to the first byte after the letterbox table
    scratch_bytes: int = 0
    max_dst_pixels: int = 0

    @property
    def table_bytes(self) -> int:
        return self.batch * _REC.itemsize

    def to_device(self, device: Optional[torch.device] = None, out: Optional[torch.Tensor] = None,
                  staging: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Copy the arena to the device (into `staging` when given: a reusable uint8 buffer at least as large) and run the
        letterbox kernel. Returns the uint8 [B, 3, H, W] RGB batch (written into `out` when given)."""
        dev = _arena_to_device(self, device, staging)
        H, W = self.out_shape
        if out is None:
            out = torch.empty((self.batch, 3, H, W), dtype=torch.uint8, device=dev.device)
        assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == (self.batch, 3, H, W)
        _launch(self, dev, LB_NCHW_U8, out.data_ptr(), 0, 0, 1.0)
        return out

    def to_space_to_depth(self, s2d, scale: float = 1.0 / 255.0, x_offset: int = 0, device: Optional[torch.device] = None,
                          staging: Optional[torch.Tensor] = None) -> None:
        """Fused form: write the stem's space-to-depth image (an `ops.ActView` over [B, H/2, Wp, 16] bf16) directly."""
        H, W = self.out_shape
        assert not s2d.x3 and s2d.c0 == 0 and s2d.cstride == 16 and (s2d.B, s2d.H) == (self.batch, H // 2)
        assert s2d.W >= W // 2 + x_offset
        dev = _arena_to_device(self, device or s2d.buf.device, staging)
        _launch(self, dev, LB_S2D_BF16, s2d.ptr(), s2d.W, x_offset, scale)


def _arena_to_device(pb: PackedBatch, device, staging: Optional[torch.Tensor]) -> torch.Tensor:
    if pb.arena.is_cuda:
        return pb.arena
    device = torch.device(device if device is not None else "cuda")
    if device.type != "cuda":
        raise RuntimeError("PackedBatch.to_device: the letterbox / collate kernel runs on a CUDA device only (no CPU fallback)")
    n = pb.arena.numel()
    if staging is None:
        if not pb.scratch_bytes:
            return pb.arena.to(device, non_blocking=True)
        staging = torch.empty(n + pb.scratch_bytes, dtype=torch.uint8, device=device)
    assert staging.is_cuda and staging.dtype == torch.uint8 and staging.numel() >= n + pb.scratch_bytes
    dst = staging[:n + pb.scratch_bytes]
    dst[:n].copy_(pb.arena, non_blocking=True)
    return dst


def _launch(pb: PackedBatch, dev_arena: torch.Tensor, kind: int, out_ptr: int, row_pixels: int, x_offset: int, scale: float) -> None:
    H, W = pb.out_shape
    base = dev_arena.data_ptr()
    assert dev_arena.numel() >= pb.host_bytes + pb.scratch_bytes, "the device arena lacks the scratch space of the decode-side resize"
    if pb.n_load:
        img_base = base + pb.table_bytes
        _lib.check(_lib.load().ay2_load_resize(img_base, img_base + pb.load_table_offset, pb.n_load, pb.max_dst_pixels,
                                               _lib.current_stream_ptr()), "ay2_load_resize")
    color = pb.color[0] | (pb.color[1] << 8) | (pb.color[2] << 16)
    _lib.check(_lib.load().ay2_letterbox_collate(base + pb.table_bytes, base, pb.batch, pb.kinds, H, W, C.c_uint32(color), kind, out_ptr,
                                                 row_pixels, x_offset, float(scale), _lib.current_stream_ptr()),
               "ay2_letterbox_collate")


def pack_batch(images: Sequence[np.ndarray], new_shape: Sequence[int], auto: bool = False, scale_fill: bool = False,
               scale_up: bool = True, stride: int = 32, color: Sequence[int] = (114, 114, 114), pin: bool = False,
               paths: Sequence[str] = (), orig_shapes: Optional[Sequence[Tuple[int, int]]] = None,
               arena: Optional[torch.Tensor] = None, img_size: Optional[int] = None, augmentation: bool = False) -> PackedBatch:
    """Host half. images: HWC BGR uint8 arrays. With `img_size=None` they are LOADED images (what `_load_image` returns,
    data_loader.py:294-350) and `orig_shapes` their (h0, w0) before that function's resize when it differs (only enters the
    `shapes` tuples). With `img_size` they are DECODED images (`cv2.imread`) and `_load_image`'s resize (:320-329: long side ->
    img_size, INTER_AREA when shrinking and `augmentation` is off, INTER_LINEAR otherwise) runs on the device too, into
    scratch space behind the uploaded bytes. The reference calls `_letterbox(img, new_shape=shape, auto=False)` (:380);
    `auto=True` is rejected here because a batch needs one shape."""
    if auto:
        raise ValueError("pack_batch: auto=True gives every image its own output shape; the collated batch needs one (the "
                         "reference's __getitem__ passes auto=False, data_loader.py:380)")
    H, W = int(new_shape[0]), int(new_shape[1])
    B = len(images)
    head = B * _REC.itemsize
    rec = np.zeros(B, dtype=_REC)
    loaded = []  # per image: (h1, w1) after the decode-side resize, or None
    for i, im in enumerate(images):
        if not (isinstance(im, np.ndarray) and im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3):
            raise TypeError(f"pack_batch: image {i} must be a uint8 HWC array with 3 channels")
        h0, w0 = im.shape[:2]
        r = img_size / max(h0, w0) if img_size is not None else 1
        loaded.append((int(h0 * r), int(w0 * r)) if r != 1 else None)
        if loaded[-1] is not None and min(loaded[-1]) < 1:
            raise ValueError(f"pack_batch: image {i} ({h0}x{w0}) vanishes at img_size {img_size}")
    n_load = sum(x is not None for x in loaded)
    load_rec = np.zeros(n_load, dtype=_LOAD_REC)
    off = (head + 15) & ~15
    load_table_offset = off - head
    off = (off + n_load * _LOAD_REC.itemsize + 15) & ~15
    src_off = []
    for im in images:
        src_off.append(off - head)
        off = (off + im.size + 15) & ~15
    host_bytes = off
    shapes, ratios, kinds, k, max_dst = [], [], 0, 0, 0
    for i, im in enumerate(images):
        h0, w0 = im.shape[:2]
        if loaded[i] is None:
            h, w, image_off = h0, w0, src_off[i]
        else:
            h, w = loaded[i]
            image_off = off - head  # in the device-only scratch space
            off = (off + 3 * h * w + 15) & ~15
            shrink = h <= h0 and w <= w0 and (h, w) != (h0, w0)
            mode = LR_AREA if shrink and not augmentation else LR_LINEAR
            load_rec[k] = (src_off[i], image_off, 1.0 / (float(w) / float(w0)), 1.0 / (float(h) / float(h0)), h0, w0, 3 * w0,
                           h, w, 3 * w, mode, 0)
            k += 1
            max_dst = max(max_dst, h * w)
        (uw, uh), ratio, (dw, dh), (top, bottom, left, right) = letterbox_geometry((h, w), (H, W), False, scale_fill, scale_up, stride)
        if (uh + top + bottom, uw + left + right) != (H, W):
            raise ValueError(f"pack_batch: image {i} ({h}x{w}) letterboxes to {uh + top + bottom}x{uw + left + right}, not {H}x{W}")
        # cv2's step per destination pixel, in its own precision and order (resize.cpp: scale = 1. / inv_scale, both double)
        sx, sy = 1.0 / (float(uw) / float(w)), 1.0 / (float(uh) / float(h))
        rec[i] = (image_off, sx, sy, h, w, 3 * w, uh, uw, top, left, 0)
        kinds |= LB_HAS_COPY if (uh, uw) == (h, w) else LB_HAS_RESIZE
        if orig_shapes is not None and img_size is None:
            h0, w0 = orig_shapes[i]
        shapes.append(((h0, w0), ((h / h0, w / w0), (dw, dh))))
        ratios.append(ratio)
    scratch_bytes = off - host_bytes
    if arena is None:
        arena = torch.empty(host_bytes, dtype=torch.uint8, pin_memory=pin)
    else:
        assert arena.dtype == torch.uint8 and not arena.is_cuda and arena.numel() >= host_bytes
        arena = arena[:host_bytes]
    host = arena.numpy()
    host[:head] = rec.view(np.uint8)
    end = head
    if n_load:
        start = head + load_table_offset
        host[end:start] = 0
        host[start:start + load_rec.nbytes] = load_rec.view(np.uint8)
        end = start + load_rec.nbytes
    for i, im in enumerate(images):
        start = head + src_off[i]
        host[end:start] = 0  # alignment gap (the arena is uninitialised memory: keep the packed bytes deterministic)
        host[start:start + im.size] = np.ascontiguousarray(im).reshape(-1)
        end = start + im.size
    host[end:host_bytes] = 0
    return PackedBatch(arena, B, (H, W), tuple(shapes), tuple(ratios), tuple(int(c) for c in color), tuple(paths), kinds,
                       host_bytes, n_load, load_table_offset, scratch_bytes, max_dst)


def collate_fn(batch: List[tuple], new_shape: Sequence[int] = (640, 640), **kw):
    """Counterpart of `LoadImages.collate_fn` (data_loader.py:461-477) for a Dataset that returns the loaded image instead of
    the letterboxed CHW tensor: batch items are (image HWC BGR uint8, path, (h0, w0)). Returns (PackedBatch, paths, shapes);
    `PackedBatch.to_device()` in the consumer (`prepare_img`) yields the tensor the reference's collate_fn returns."""
    imgs, paths, orig = zip(*batch) if batch else ((), (), ())
    pb = pack_batch(list(imgs), new_shape, paths=paths, orig_shapes=orig, **kw)
    return pb, tuple(paths), pb.shapes


def collate_labels(labels: Sequence[torch.Tensor], device: Optional[torch.device] = None) -> torch.Tensor:
    """`LoadImagesAndLabels.collate_fn`'s label half (data_loader.py:905-909): rows of all images concatenated, column 0 set
    to the image's index in the batch -- one copy + one launch instead of a Python loop of in-place writes."""
    device = torch.device(device if device is not None else "cuda")
    counts = [int(l.shape[0]) for l in labels]
    total = sum(counts)
    if total == 0:
        return torch.zeros((0, 6), dtype=torch.float32, device=device)
    flat = torch.cat([torch.as_tensor(l, dtype=torch.float32).reshape(-1, 6) for l in labels], 0).to(device, non_blocking=True)
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)).to(device, non_blocking=True)
    _lib.check(_lib.load().ay2_collate_labels(flat.data_ptr(), offsets.data_ptr(), len(labels), total, _lib.current_stream_ptr()),
               "ay2_collate_labels")
    return flat
