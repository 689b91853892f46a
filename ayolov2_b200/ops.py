"""Thin tensor-level wrappers over the C-ABI (one function / class per libay2 entry point).

Everything here takes CUDA tensors that the caller owns; layouts are NHWC bf16 with explicit channel
strides (a "view" of a wider buffer is expressed as (tensor, channel_offset, channels)).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_SILU, ChainDesc, ConvDesc, HeadLevels, NmsParams


@dataclass
class ActView:
    """A channel slice [c0, c0+c) of an NHWC bf16 buffer of shape [B, H, W, Cs].

    x3 = split-precision layout (include/ay2.h, ay2_conv_desc::x3): c0 / c stay LOGICAL channel numbers; every producer's
    output segment of w channels occupies 3 w physical channels [hi | lo | hi], so logical offset c0 sits at physical 3 c0
    (views are only ever cut at segment boundaries) and the buffer's last dimension is 3x the logical width."""

    buf: torch.Tensor
    c0: int
    c: int
    x3: bool = False

    @property
    def B(self) -> int:
        return self.buf.shape[0]

    @property
    def H(self) -> int:
        return self.buf.shape[1]

    @property
    def W(self) -> int:
        return self.buf.shape[2]

    @property
    def cstride(self) -> int:
        return self.buf.shape[3]

    @property
    def pc(self) -> int:
        """physical channels of the view"""
        return 3 * self.c if self.x3 else self.c

    def ptr(self) -> int:
        return self.buf.data_ptr() + 2 * (3 * self.c0 if self.x3 else self.c0)

    def tensor(self) -> torch.Tensor:
        assert not self.x3, "a split-precision view has no plain tensor form (use value_f32 on a single segment)"
        return self.buf[..., self.c0:self.c0 + self.c]

    def value_f32(self) -> torch.Tensor:
        """fp32 [B, H, W, c] of a view that is ONE split-precision segment: hi + lo."""
        assert self.x3
        p0 = 3 * self.c0
        return self.buf[..., p0:p0 + self.c].float() + self.buf[..., p0 + self.c:p0 + 2 * self.c].float()

    def slice(self, c0: int, c: int) -> "ActView":
        assert 0 <= c0 and c0 + c <= self.c
        return ActView(self.buf, self.c0 + c0, c, self.x3)


def new_act(B: int, H: int, W: int, C_: int, device="cuda", x3: bool = False) -> ActView:
    assert C_ % 8 == 0
    return ActView(torch.empty((B, H, W, 3 * C_ if x3 else C_), dtype=torch.bfloat16, device=device), 0, C_, x3)


def conv_block_n(cout: int) -> int:
    return int(_lib.load().ay2_conv_block_n(cout))


def pack_conv_weight(w: torch.Tensor, bias: Optional[torch.Tensor], bn: Optional[Tuple[torch.Tensor, ...]] = None,
                     eps: float = 1e-3) -> Tuple[torch.Tensor, torch.Tensor]:
    """OIHW fp32 weight (+ optional BatchNorm (gamma, beta, mean, var)) -> K-major bf16 [Cout_pad, KH*KW*Cin]
    and fp32 bias [Cout_pad] with the BN folded in (what kindle's YOLOModel.fuse() does, val.py:331)."""
    w = w.detach().float()
    cout, cin, kh, kw = w.shape
    if bn is not None:
        gamma, beta, mean, var = [t.detach().float() for t in bn]
        scale = gamma / torch.sqrt(var + eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = beta - mean * scale
        if bias is not None:
            b = b + bias.detach().float() * scale
    else:
        b = bias.detach().float() if bias is not None else torch.zeros(cout, device=w.device)
    bn_tile = conv_block_n(cout)
    cout_pad = (cout + bn_tile - 1) // bn_tile * bn_tile
    wk = w.permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)
    wp = torch.zeros((cout_pad, kh * kw * cin), dtype=torch.bfloat16, device=w.device)
    wp[:cout] = wk.to(torch.bfloat16)
    bp = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    bp[:cout] = b
    return wp.contiguous(), bp.contiguous()


def pack_conv_weight_x3(w: torch.Tensor, bias: Optional[torch.Tensor], bn: Optional[Tuple[torch.Tensor, ...]], eps: float,
                        segs: Sequence[Tuple[int, int]]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Split-precision weights: the BN-folded fp32 weight w = w_hi + w_lo (both bf16) laid out along K to meet the input's
    segment planes: for every input segment (offset, width): [w_hi | w_hi | w_lo] of its channels. segs must tile cin."""
    w = w.detach().float()
    cout, cin, kh, kw = w.shape
    if bn is not None:
        gamma, beta, mean, var = [t.detach().float() for t in bn]
        scale = gamma / torch.sqrt(var + eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = beta - mean * scale
        if bias is not None:
            b = b + bias.detach().float() * scale
    else:
        b = bias.detach().float() if bias is not None else torch.zeros(cout, device=w.device)
    assert sum(wd for _, wd in segs) == cin and [o for o, _ in segs] == [sum(wd for _, wd in segs[:i]) for i in range(len(segs))], segs
    hi = w.to(torch.bfloat16).float()
    lo = (w - hi).to(torch.bfloat16).float()
    parts = []
    for off, wd in segs:
        parts += [hi[:, off:off + wd], hi[:, off:off + wd], lo[:, off:off + wd]]
    w3 = torch.cat(parts, 1)
    bn_tile = conv_block_n(cout)
    cout_pad = (cout + bn_tile - 1) // bn_tile * bn_tile
    wp = torch.zeros((cout_pad, kh * kw * 3 * cin), dtype=torch.bfloat16, device=w.device)
    wp[:cout] = w3.permute(0, 2, 3, 1).reshape(cout, -1).to(torch.bfloat16)
    bp = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    bp[:cout] = b
    return wp.contiguous(), bp.contiguous()


class ConvPlan:
    """ay2_conv_plan: fused conv + bias + act (+ residual) between two ActViews. Keeps its tensors alive."""

    def __init__(self, x: ActView, y: ActView, w_packed: torch.Tensor, bias: torch.Tensor, kh: int, kw: int,
                 stride: int, pad: int, act: int, residual: Optional[ActView] = None, pad_w: int = -1,
                 window: Optional[Tuple[int, int, int, int]] = None, out_sub: Optional[Tuple[int, int]] = None,
                 x2: Optional[ActView] = None, stride_w: int = 0):
        """`window` = (cin, in_w, pix_stride, row_pixels): read `x.buf` as overlapping windows of `cin` channels
        starting at every physical pixel (the packed 16-channel stem); otherwise the input is the ActView `x`.
        `out_sub` = (py, px): write (and read the residual from) the (row, column) parity sub-grid of `y` -- the
        output is then y.H/2 x y.W/2 pixels (data gradient of a stride-2 convolution).
        `x2`: second input view; the conv consumes torch.cat([x, x2], channel) without the concatenation existing.
        `stride_w`: horizontal stride when it differs from `stride` (the pixel-pair form: stride 2 over rows, 1 over pairs)."""
        lib = _lib.load()
        cout_pad, ktot = w_packed.shape
        cin = window[0] if window else x.pc + (x2.pc if x2 is not None else 0)
        assert ktot == kh * kw * cin, (ktot, kh, kw, cin)
        x3 = y.x3
        assert x.x3 == x3 and (x2 is None or x2.x3 == x3) and (residual is None or residual.x3 == x3)
        assert w_packed.dtype == torch.bfloat16 and bias.dtype == torch.float32 and bias.numel() == cout_pad
        assert x.buf.is_cuda and y.buf.is_cuda and w_packed.is_cuda and bias.is_cuda
        d = ConvDesc()
        d.batch = x.B
        d.in_h, d.in_w, d.cin, d.in_cstride = x.H, x.W, cin, x.cstride
        d.pad_w = pad_w
        if x2 is not None:
            assert window is None and (x2.B, x2.H, x2.W) == (x.B, x.H, x.W) and x2.buf.is_cuda
            d.cin_split, d.in2_cstride, d.in2 = x.pc, x2.cstride, x2.ptr()
        if window:
            d.cin, d.in_w, d.in_pix_stride, d.in_row_pixels = window
        d.out_h, d.out_w, d.cout, d.out_cstride = y.H, y.W, y.c, y.cstride
        y_ptr = y.ptr()
        r_ptr = residual.ptr() if residual is not None else None
        if out_sub is not None:
            py, px = out_sub
            assert y.H % 2 == 0 and y.W % 2 == 0
            d.out_h, d.out_w = y.H // 2, y.W // 2
            d.out_pix_stride, d.out_row_pixels = 2 * y.cstride, y.W
            y_ptr += 2 * (py * y.W + px) * y.cstride
            if residual is not None:
                assert (residual.H, residual.W) == (y.H, y.W)
                r_ptr += 2 * (py * residual.W + px) * residual.cstride
        d.kh, d.kw, d.stride, d.pad, d.act = kh, kw, stride, pad, act
        d.stride_w = stride_w
        d.res_cstride = residual.cstride if residual is not None else 0
        d.cout_pad = cout_pad
        d.x3 = int(x3)
        assert not (x3 and out_sub is not None)
        self.desc = d
        self.x, self.y, self.w, self.b, self.res, self.x2 = x, y, w_packed, bias, residual, x2
        h = C.c_void_p()
        _lib.check(lib.ay2_conv_plan_create(C.byref(d), x.ptr(), w_packed.data_ptr(), bias.data_ptr(), r_ptr, y_ptr,
                                            C.byref(h)), "ay2_conv_plan_create")
        self._h = h
        self._lib = lib
        self.flops = float(lib.ay2_conv_plan_flops(h))
        info = (C.c_int32 * 4)()
        lib.ay2_conv_plan_set_debug(h, None, info)
        self.halo = info[2] < 0  # 3x3 / s1 halo kernel (conv_halo_kernel) instead of conv_tc_kernel
        info8 = (C.c_int32 * 8)()
        lib.ay2_conv_plan_info(h, info8)
        self.pair = bool(info8[5])  # CTA-pair form (tcgen05 cta_group::2)

    def run(self, stream: Optional[int] = None) -> None:
        _lib.check(self._lib.ay2_conv_plan_run(self._h, stream if stream is not None else _lib.current_stream_ptr()),
                   "ay2_conv_plan_run")

    def set_head_candidates(self, ws: Optional["NmsWorkspace"], na: int = 0, row_off: int = 0,
                            class_mask: Optional[torch.Tensor] = None) -> None:
        """Detect-head plan: also append this level's NMS candidates to `ws` on every run (ws.p carries conf_thres,
        multi_label, max_candidates). ws=None switches it off."""
        if ws is None:
            _lib.check(self._lib.ay2_conv_plan_set_head_candidates(self._h, None, 0, 0, None, None, 0),
                       "ay2_conv_plan_set_head_candidates")
            self._cand_keep = None
            return
        self._cand_keep = (ws, class_mask)
        _lib.check(self._lib.ay2_conv_plan_set_head_candidates(self._h, C.byref(ws.p), na, row_off, _lib.ptr(class_mask),
                                                               ws.ws.data_ptr(), ws.ws.numel()),
                   "ay2_conv_plan_set_head_candidates")

    def run_reference_simt(self) -> None:
        """Same math on CUDA cores (test infrastructure)."""
        _lib.check(self._lib.ay2_conv_reference_simt(C.byref(self.desc), self.x.ptr(), self.w.data_ptr(),
                                                     self.b.data_ptr(), self.res.ptr() if self.res is not None else None,
                                                     self.y.ptr(), _lib.current_stream_ptr()), "ay2_conv_reference_simt")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.ay2_conv_plan_destroy(h)
            self._h = None


def pack_chain_weight(w: torch.Tensor, bias: Optional[torch.Tensor], bn: Optional[Tuple[torch.Tensor, ...]] = None,
                      eps: float = 1e-3, cin_pad: Optional[int] = None, cout_pad: Optional[int] = None):
    """One link of a fused chain: OIHW fp32 (+ optional BN fold) -> bf16 K-major [cout_pad][kh*kw*cin_pad] and fp32 bias
    [cout_pad]; input / output channels are zero-padded to multiples of 16 (a padded output channel is act(0) = 0 and
    meets zero weights in the next link)."""
    w = w.detach().float()
    cout, cin, kh, kw = w.shape
    if bn is not None:
        gamma, beta, mean, var = [t.detach().float() for t in bn]
        scale = gamma / torch.sqrt(var + eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = beta - mean * scale
        if bias is not None:
            b = b + bias.detach().float() * scale
    else:
        b = bias.detach().float() if bias is not None else torch.zeros(cout, device=w.device)
    cin_pad = cin_pad or (cin + 15) // 16 * 16
    cout_pad = cout_pad or (cout + 15) // 16 * 16
    wp = torch.zeros((cout_pad, kh, kw, cin_pad), dtype=torch.float32, device=w.device)
    wp[:cout, :, :, :cin] = w.permute(0, 2, 3, 1)
    bp = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    bp[:cout] = b
    return wp.reshape(cout_pad, kh * kw * cin_pad).to(torch.bfloat16).contiguous(), bp


def chain_desc(x: ActView, c1: int, c2: int, c3: int, act1: int, act2: int, act3: int, out_cstride: int,
               res_cstride: int, k: int = 3, stride: int = 1, pad: int = 1) -> ChainDesc:
    d = ChainDesc()
    d.batch, d.in_h, d.in_w, d.cin, d.in_cstride = x.B, x.H, x.W, x.c, x.cstride
    d.c1, d.act1, d.c2, d.act2, d.c3, d.act3 = c1, act1, c2, act2, c3, act3
    d.kh = d.kw = k
    d.stride, d.pad = stride, pad
    d.out_cstride, d.res_cstride = out_cstride, res_cstride
    return d


def chain_supported(d: ChainDesc) -> bool:
    return bool(_lib.load().ay2_chain_supported(C.byref(d)))


class ChainPlan:
    """ay2_chain_plan: 1x1 -> 3x3 (-> 1x1) fused in one kernel (Tucker-2 chain / Bottleneck). Keeps its tensors alive.
    links = [(w_packed, bias)] * 2 or 3 from pack_chain_weight; acts = activation code per link."""

    def __init__(self, x: ActView, y: ActView, links, acts, residual: Optional[ActView] = None):
        lib = _lib.load()
        assert len(links) in (2, 3) and len(acts) == len(links)
        (w1, b1), (w2, b2) = links[0], links[1]
        w3, b3 = links[2] if len(links) == 3 else (None, None)
        c1, c2 = w1.shape[0], w2.shape[0]
        c3 = w3.shape[0] if w3 is not None else 0
        assert w1.shape[1] == x.c and w2.shape[1] == 9 * c1 and (w3 is None or w3.shape[1] == c2)
        assert y.c == (c3 or c2) and (y.H, y.W) == (x.H, x.W)
        d = chain_desc(x, c1, c2, c3, acts[0], acts[1], acts[2] if c3 else ACT_NONE, y.cstride,
                       residual.cstride if residual is not None else 0)
        bias = torch.cat([b1, b2] + ([b3] if b3 is not None else [])).float().contiguous()
        self.desc = d
        self.keep = (x, y, w1, w2, w3, bias, residual)
        h = C.c_void_p()
        _lib.check(lib.ay2_chain_plan_create(C.byref(d), x.ptr(), w1.data_ptr(), w2.data_ptr(), _lib.ptr(w3), bias.data_ptr(),
                                             residual.ptr() if residual is not None else None, y.ptr(), C.byref(h)),
                   "ay2_chain_plan_create")
        self._h, self._lib = h, lib
        self.flops = float(lib.ay2_chain_plan_flops(h))

    def run(self, stream: Optional[int] = None) -> None:
        _lib.check(self._lib.ay2_chain_plan_run(self._h, stream if stream is not None else _lib.current_stream_ptr()),
                   "ay2_chain_plan_run")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.ay2_chain_plan_destroy(h)
            self._h = None


def space_to_depth(img: torch.Tensor, out: ActView, scale: float, x_offset: int = 0) -> None:
    """img: NCHW uint8/fp32 [B,3,H,W] -> out [B,H/2,Wp,16] with logical pixel x at physical column x + x_offset
    (Wp >= W/2 + x_offset; untouched columns keep whatever the caller put there, i.e. zeros = conv padding)."""
    assert img.is_cuda and img.is_contiguous() and img.shape[1] == 3
    B, _, H, W = img.shape
    assert out.c0 == 0 and out.cstride == 16 and (out.B, out.H) == (B, H // 2) and out.W >= W // 2 + x_offset
    dt = {torch.uint8: _lib.DT_U8, torch.float32: _lib.DT_F32}[img.dtype]
    _lib.check(_lib.load().ay2_space_to_depth(img.data_ptr(), dt, B, H, W, float(scale), out.ptr(), out.W, x_offset,
                                              _lib.current_stream_ptr()), "ay2_space_to_depth")


def resize_bilinear(img: torch.Tensor, out: torch.Tensor, pre_scale: float = 1.0) -> None:
    """out (fp32 NCHW, any size) = F.interpolate(img * pre_scale, out.shape[2:], mode="bilinear", align_corners=False);
    img NCHW uint8 / fp32. `YoloTrainer.multi_scale` + `prepare_img` in one pass (yolo_trainer.py:223-248)."""
    assert img.is_cuda and out.is_cuda and img.is_contiguous() and out.is_contiguous() and out.dtype == torch.float32
    assert img.dim() == 4 and img.shape[1] == 3 and out.shape[:2] == img.shape[:2]
    dt = {torch.uint8: _lib.DT_U8, torch.float32: _lib.DT_F32}[img.dtype]
    _lib.check(_lib.load().ay2_resize_bilinear(img.data_ptr(), dt, img.shape[0], img.shape[2], img.shape[3], float(pre_scale),
                                               out.data_ptr(), out.shape[2], out.shape[3], _lib.current_stream_ptr()),
               "ay2_resize_bilinear")


def sppf_pool(x: ActView, o1: ActView, o2: ActView, o3: ActView, ks: Sequence[int]) -> None:
    assert x.cstride == o1.cstride == o2.cstride == o3.cstride and x.buf is o1.buf
    _lib.check(_lib.load().ay2_sppf_pool(x.ptr(), x.B, x.H, x.W, x.c, x.cstride, ks[0], ks[1], ks[2], o1.ptr(),
                                         o2.ptr(), o3.ptr(), _lib.current_stream_ptr()), "ay2_sppf_pool")


def upsample2x(x: ActView, y: ActView) -> None:
    assert (y.H, y.W, y.c) == (2 * x.H, 2 * x.W, x.c) and x.x3 == y.x3
    _lib.check(_lib.load().ay2_upsample2x(x.ptr(), x.B, x.H, x.W, x.pc, x.cstride, y.ptr(), y.cstride,
                                          _lib.current_stream_ptr()), "ay2_upsample2x")


def space_to_depth_x3(img: torch.Tensor, out: ActView, divisor: float, x_offset: int = 0) -> None:
    """Split-precision space-to-depth: out.buf [B, H/2, Wp, 48] = planes [hi16 | lo16 | hi16] of img / divisor."""
    assert img.is_cuda and img.is_contiguous() and img.shape[1] == 3 and out.x3 and out.buf.shape[3] == 48
    B, _, H, W = img.shape
    dt = {torch.uint8: _lib.DT_U8, torch.float32: _lib.DT_F32}[img.dtype]
    _lib.check(_lib.load().ay2_space_to_depth_x3(img.data_ptr(), dt, B, H, W, float(divisor), out.buf.data_ptr(), out.W, x_offset,
                                                 _lib.current_stream_ptr()), "ay2_space_to_depth_x3")


def sppf_pool_x3(x: ActView, o1: ActView, o2: ActView, o3: ActView, ks: Sequence[int]) -> None:
    assert x.x3 and x.cstride == o1.cstride == o2.cstride == o3.cstride and x.buf is o1.buf
    _lib.check(_lib.load().ay2_sppf_pool_x3(x.ptr(), x.B, x.H, x.W, x.c, x.cstride, ks[0], ks[1], ks[2], o1.ptr(), o2.ptr(),
                                            o3.ptr(), _lib.current_stream_ptr()), "ay2_sppf_pool_x3")


def head_decode2(logits: ActView, na: int, no: int, stride_px: float, anchor_wh_px: torch.Tensor, pred: torch.Tensor,
                 row_offset: int, raw: Optional[torch.Tensor], xyxy: bool = False) -> None:
    """General head decode: split-precision logits (hi + lo planes, exact sigmoid) and / or x1 y1 x2 y2 boxes."""
    assert logits.c0 == 0 and pred.dtype == torch.float32 and pred.is_contiguous() and pred.shape[2] == no
    flags = (1 if xyxy else 0) | (2 if logits.x3 else 0)
    _lib.check(_lib.load().ay2_head_decode2(logits.ptr(), logits.c if logits.x3 else 0, logits.B, logits.H, logits.W,
                                            logits.cstride, na, no, float(stride_px), anchor_wh_px.data_ptr(), flags,
                                            pred.data_ptr(), pred.shape[1], row_offset, _lib.ptr(raw),
                                            _lib.current_stream_ptr()), "ay2_head_decode2")


def head_decode(logits: ActView, na: int, no: int, stride_px: float, anchor_wh_px: torch.Tensor, pred: torch.Tensor,
                row_offset: int, raw: Optional[torch.Tensor]) -> None:
    assert logits.c0 == 0 and pred.dtype == torch.float32 and pred.is_contiguous() and pred.shape[2] == no
    assert anchor_wh_px.dtype == torch.float32 and anchor_wh_px.numel() == na * 2 and anchor_wh_px.is_cuda
    _lib.check(_lib.load().ay2_head_decode(logits.ptr(), logits.B, logits.H, logits.W, logits.cstride, na, no,
                                           float(stride_px), anchor_wh_px.data_ptr(), pred.data_ptr(), pred.shape[1],
                                           row_offset, _lib.ptr(raw), _lib.current_stream_ptr()), "ay2_head_decode")


class NmsWorkspace:
    """Pre-allocated buffers for ay2_nms_batched on a fixed (batch, n, no) problem."""

    def __init__(self, batch: int, n: int, no: int, max_det: int = 300, multi_label: bool = False,
                 max_candidates: Optional[int] = None, device="cuda"):
        nc = no - 5
        if max_candidates is None:
            max_candidates = n if not multi_label else min(n * nc, max(n, 131072))
        self.p = NmsParams()
        self.p.batch, self.p.n, self.p.no = batch, n, no
        self.p.max_det = max_det
        self.p.max_candidates = max_candidates
        self.p.multi_label = int(multi_label)
        lib = _lib.load()
        nbytes = lib.ay2_nms_workspace_bytes(C.byref(self.p))
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.out = torch.zeros((batch, max_det, 6), dtype=torch.float32, device=device)
        self.count = torch.zeros(batch, dtype=torch.int32, device=device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=device)

    def run(self, pred: torch.Tensor, conf_thres: float, iou_thres: float, agnostic: bool = False,
            class_mask: Optional[torch.Tensor] = None, max_nms: int = 30000, max_wh: float = 4096.0) -> None:
        p = self.p
        assert pred.is_cuda and pred.dtype == torch.float32 and pred.is_contiguous()
        assert tuple(pred.shape) == (p.batch, p.n, p.no), (pred.shape, (p.batch, p.n, p.no))
        p.conf_thres, p.iou_thres = float(conf_thres), float(iou_thres)
        p.agnostic, p.max_nms, p.max_wh = int(agnostic), int(max_nms), float(max_wh)
        _lib.check(_lib.load().ay2_nms_batched(pred.data_ptr(), C.byref(p), _lib.ptr(class_mask), self.ws.data_ptr(),
                                               self.ws.numel(), self.out.data_ptr(), self.count.data_ptr(),
                                               self.overflow.data_ptr(), _lib.current_stream_ptr()), "ay2_nms_batched")

    def run_logits(self, levels: "HeadLevels", keep, conf_thres: float, iou_thres: float, agnostic: bool = False,
                   class_mask: Optional[torch.Tensor] = None, max_nms: int = 30000, max_wh: float = 4096.0) -> None:
        """Fused head: NMS straight from the bf16 head logits (ay2_nms_from_logits). `keep` = tensors to keep alive."""
        p = self.p
        p.conf_thres, p.iou_thres = float(conf_thres), float(iou_thres)
        p.agnostic, p.max_nms, p.max_wh = int(agnostic), int(max_nms), float(max_wh)
        self._keepalive = keep
        _lib.check(_lib.load().ay2_nms_from_logits(C.byref(levels), C.byref(p), _lib.ptr(class_mask), self.ws.data_ptr(),
                                                   self.ws.numel(), self.out.data_ptr(), self.count.data_ptr(),
                                                   self.overflow.data_ptr(), _lib.current_stream_ptr()),
                   "ay2_nms_from_logits")


    def begin_candidates(self) -> None:
        """Zero the per-image candidate counters; the detect-head convolutions that follow append candidates
        (ConvPlan.set_head_candidates), run_candidates() finishes the step."""
        _lib.check(_lib.load().ay2_nms_candidates_begin(C.byref(self.p), self.ws.data_ptr(), self.ws.numel(),
                                                        _lib.current_stream_ptr()), "ay2_nms_candidates_begin")

    def run_candidates(self, levels: "HeadLevels", keep, iou_thres: float, agnostic: bool = False, max_nms: int = 30000,
                       max_wh: float = 4096.0) -> None:
        """Sort + suppression + output over the candidates the head convolutions produced (ay2_nms_from_candidates).
        conf_thres / multi_label are the ones in self.p when the plans were armed."""
        p = self.p
        p.iou_thres = float(iou_thres)
        p.agnostic, p.max_nms, p.max_wh = int(agnostic), int(max_nms), float(max_wh)
        self._keepalive = keep
        _lib.check(_lib.load().ay2_nms_from_candidates(C.byref(levels), C.byref(p), self.ws.data_ptr(), self.ws.numel(),
                                                       self.out.data_ptr(), self.count.data_ptr(), self.overflow.data_ptr(),
                                                       _lib.current_stream_ptr()), "ay2_nms_from_candidates")


def make_head_levels(logits: Sequence[ActView], na: int, strides: Sequence[float], anchors_px: Sequence[Sequence[Sequence[float]]]) -> HeadLevels:
    hl = HeadLevels()
    hl.nl, hl.na = len(logits), na
    for i, lv in enumerate(logits):
        assert lv.c0 == 0
        hl.logits[i] = lv.ptr()
        hl.ny[i], hl.nx[i], hl.cstride[i] = lv.H, lv.W, lv.cstride
        hl.stride_px[i] = float(strides[i])
        for a in range(na):
            hl.anchor_px[i][a][0] = float(anchors_px[i][a][0])
            hl.anchor_px[i][a][1] = float(anchors_px[i][a][1])
    return hl


# -------------------------------------------------------------------------------------------------
# training-step wrappers (include/ay2.h "Training step pieces")
# -------------------------------------------------------------------------------------------------
def _npix(v: ActView) -> int:
    return v.B * v.H * v.W


def _sync_world(sync: bool) -> int:
    import torch.distributed as dist

    return dist.get_world_size() if (sync and dist.is_available() and dist.is_initialized()) else 1


def bn_batch_stats(z: ActView, eps: float, momentum: float, running_mean: Optional[torch.Tensor],
                   running_var: Optional[torch.Tensor], scratch: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor,
                   sync: bool = False, zeroed: bool = False) -> None:
    """scratch: double[2*c] (zeroed here unless the caller already did: `zeroed`); mean / invstd: fp32 [c] outputs; running
    stats updated in place. sync (SyncBatchNorm): the per-channel sums are all-reduced, the statistics are those of the
    cross-rank batch."""
    lib = _lib.load()
    c = z.c
    if not zeroed:
        scratch.zero_()
    st = _lib.current_stream_ptr()
    _lib.check(lib.ay2_bn_stats(z.ptr(), _npix(z), c, z.cstride, scratch.data_ptr(), scratch.data_ptr() + 8 * c, st), "ay2_bn_stats")
    world = _sync_world(sync)
    if world > 1:
        torch.distributed.all_reduce(scratch)
    _lib.check(lib.ay2_bn_finalize(scratch.data_ptr(), scratch.data_ptr() + 8 * c, _npix(z) * world, c, float(eps), float(momentum),
                                   _lib.ptr(running_mean), _lib.ptr(running_var), mean.data_ptr(), invstd.data_ptr(), st),
               "ay2_bn_finalize")


def bn_act_fwd(z: ActView, mean, invstd, gamma, beta, act: int, y: ActView, residual: Optional[ActView] = None) -> None:
    _lib.check(_lib.load().ay2_bn_act_fwd(z.ptr(), _npix(z), z.c, z.cstride, mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                          beta.data_ptr(), act, y.ptr(), y.cstride,
                                          residual.ptr() if residual is not None else None,
                                          residual.cstride if residual is not None else 0, _lib.current_stream_ptr()),
               "ay2_bn_act_fwd")


def bn_act_bwd(dy: ActView, z: ActView, mean, invstd, gamma, beta, act: int, sums: torch.Tensor, dz: ActView,
               sync: bool = False, grad_beta: Optional[torch.Tensor] = None, grad_gamma: Optional[torch.Tensor] = None,
               zeroed: bool = False) -> bool:
    """sums: double[2*c]; afterwards sums[:c] = d beta, sums[c:] = d gamma (under `sync` with more than one rank: of the
    cross-rank batch -- SyncBatchNorm's backward; DDP then averages them like every other gradient, so divide by world).
    grad_beta / grad_gamma: fp32 [c] accumulators (slices of the flat gradient) the kernel adds the two sums to; returns
    True when it did (single-rank statistics), False when the caller has to add `sums` itself. zeroed: `sums` is already
    zero (honoured on the single-rank path only; the SyncBatchNorm path clears it itself)."""
    c = z.c
    world = _sync_world(sync)
    if world > 1:
        lib, st = _lib.load(), _lib.current_stream_ptr()
        args = (dy.ptr(), dy.cstride, z.ptr(), z.cstride, _npix(z), c, mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                beta.data_ptr(), act, sums.data_ptr(), sums.data_ptr() + 8 * c, dz.ptr(), dz.cstride)
        _lib.check(lib.ay2_bn_act_bwd_phase(*args, 1, _npix(z) * world, st), "ay2_bn_act_bwd_phase")
        torch.distributed.all_reduce(sums)
        _lib.check(lib.ay2_bn_act_bwd_phase(*args, 2, _npix(z) * world, st), "ay2_bn_act_bwd_phase")
        sums.div_(world)  # this rank's share: the gradient all-reduce (sum, then / world) restores the total
        return False
    if grad_beta is not None and grad_gamma is not None:
        assert grad_beta.dtype == grad_gamma.dtype == torch.float32 and grad_beta.is_contiguous() and grad_gamma.is_contiguous()
        assert grad_beta.numel() == c and grad_gamma.numel() == c
        _lib.check(_lib.load().ay2_bn_act_bwd_grads(dy.ptr(), dy.cstride, z.ptr(), z.cstride, _npix(z), c, mean.data_ptr(),
                                                    invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act, sums.data_ptr(),
                                                    sums.data_ptr() + 8 * c, dz.ptr(), dz.cstride, grad_beta.data_ptr(),
                                                    grad_gamma.data_ptr(), int(zeroed), _lib.current_stream_ptr()),
                   "ay2_bn_act_bwd_grads")
        return True
    _lib.check(_lib.load().ay2_bn_act_bwd(dy.ptr(), dy.cstride, z.ptr(), z.cstride, _npix(z), c, mean.data_ptr(), invstd.data_ptr(),
                                          gamma.data_ptr(), beta.data_ptr(), act, sums.data_ptr(), sums.data_ptr() + 8 * c,
                                          dz.ptr(), dz.cstride, _lib.current_stream_ptr()), "ay2_bn_act_bwd")
    return False


def conv_wgrad(x: ActView, dz: ActView, dw: torch.Tensor, kh: int, kw: int, stride: int, pad: int,
               in_row_pixels: int = 0, x_ptr: Optional[int] = None, in_w: Optional[int] = None) -> None:
    """dw: fp32 [cout, kh*kw*cin] accumulated (+=)."""
    d = ConvDesc()
    d.batch = x.B
    d.in_h, d.in_w, d.cin, d.in_cstride = x.H, (in_w if in_w is not None else x.W), x.c, x.cstride
    d.out_h, d.out_w, d.cout, d.out_cstride = dz.H, dz.W, dz.c, dz.cstride
    d.kh, d.kw, d.stride, d.pad = kh, kw, stride, pad
    d.pad_w = -1
    d.in_row_pixels = in_row_pixels
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == dz.c * kh * kw * x.c
    _lib.check(_lib.load().ay2_conv_wgrad(C.byref(d), x_ptr if x_ptr is not None else x.ptr(), dz.ptr(), dw.data_ptr(),
                                          _lib.current_stream_ptr()), "ay2_conv_wgrad")


def add_slices(src: ActView, dst: ActView, accumulate: bool) -> None:
    assert (src.B, src.H, src.W, src.c) == (dst.B, dst.H, dst.W, dst.c)
    _lib.check(_lib.load().ay2_add_slices(src.ptr(), src.cstride, dst.ptr(), dst.cstride, _npix(src), src.c, int(accumulate),
                                          _lib.current_stream_ptr()), "ay2_add_slices")


def upsample2x_bwd(dy: ActView, dx: ActView, accumulate: bool) -> None:
    assert (dy.H, dy.W, dy.c) == (2 * dx.H, 2 * dx.W, dx.c)
    _lib.check(_lib.load().ay2_upsample2x_bwd(dy.ptr(), dy.cstride, dx.B, dx.H, dx.W, dx.c, dx.ptr(), dx.cstride,
                                              int(accumulate), _lib.current_stream_ptr()), "ay2_upsample2x_bwd")


def maxpool_bwd(x: ActView, dy: ActView, k: int, dx: ActView, accumulate: bool) -> None:
    _lib.check(_lib.load().ay2_maxpool_bwd(x.ptr(), x.cstride, dy.ptr(), dy.cstride, x.B, x.H, x.W, x.c, k, dx.ptr(), dx.cstride,
                                           int(accumulate), _lib.current_stream_ptr()), "ay2_maxpool_bwd")


def head_grad_to_nhwc(grad: torch.Tensor, out: ActView) -> None:
    B, na, ny, nx, no = grad.shape
    assert grad.dtype == torch.float32 and grad.is_contiguous() and out.c0 == 0
    _lib.check(_lib.load().ay2_head_grad_to_nhwc(grad.data_ptr(), B, na, ny, nx, no, out.ptr(), out.cstride,
                                                 _lib.current_stream_ptr()), "ay2_head_grad_to_nhwc")


def head_logits_to_train(logits: ActView, na: int, no: int, out: torch.Tensor) -> None:
    assert out.dtype == torch.float32 and out.is_contiguous() and logits.c0 == 0
    _lib.check(_lib.load().ay2_head_logits_to_train(logits.ptr(), logits.cstride, logits.B, na, logits.H, logits.W, no,
                                                    out.data_ptr(), _lib.current_stream_ptr()), "ay2_head_logits_to_train")


def channel_sum(g: ActView, out: torch.Tensor) -> None:
    """out: double[c], accumulated."""
    _lib.check(_lib.load().ay2_channel_sum(g.ptr(), _npix(g), g.c, g.cstride, out.data_ptr(), _lib.current_stream_ptr()),
               "ay2_channel_sum")


def sgd_ema_step(param: torch.Tensor, grad: torch.Tensor, mom: torch.Tensor, ema: Optional[torch.Tensor], lr: float,
                 momentum: float, weight_decay: float, nesterov: bool, ema_decay: float,
                 inv_scale: Optional[torch.Tensor] = None) -> None:
    assert param.dtype == grad.dtype == mom.dtype == torch.float32 and param.is_contiguous() and grad.is_contiguous()
    _lib.check(_lib.load().ay2_sgd_ema_step(param.data_ptr(), grad.data_ptr(), mom.data_ptr(), _lib.ptr(ema), param.numel(),
                                            float(lr), float(momentum), float(weight_decay), int(nesterov), float(ema_decay),
                                            _lib.ptr(inv_scale), _lib.current_stream_ptr()), "ay2_sgd_ema_step")
    # the kernel writes through raw pointers: tell autograd / the compiled-engine cache that these tensors changed
    torch._C._increment_version([t for t in (param, mom, ema) if t is not None])  # takes an ITERABLE of tensors


def sgd_ema_step_groups(param: torch.Tensor, grad: torch.Tensor, mom: torch.Tensor, ema: Optional[torch.Tensor],
                        group: torch.Tensor, lr: Sequence[float], wd: Sequence[float], momentum: float, nesterov: bool,
                        ema_decay: float, grad_scale: float = 1.0) -> None:
    """Fused SGD + EMA over a flat parameter buffer with per-element groups (group: uint8 [n], values 0..3)."""
    assert param.dtype == grad.dtype == mom.dtype == torch.float32 and group.dtype == torch.uint8
    assert param.is_contiguous() and grad.is_contiguous() and group.numel() == param.numel()
    lr4 = (C.c_float * 4)(*(list(lr) + [0.0] * 4)[:4])
    wd4 = (C.c_float * 4)(*(list(wd) + [0.0] * 4)[:4])
    _lib.check(_lib.load().ay2_sgd_ema_step_groups(param.data_ptr(), grad.data_ptr(), mom.data_ptr(), _lib.ptr(ema), group.data_ptr(),
                                                   param.numel(), lr4, wd4, float(momentum), int(nesterov), float(ema_decay),
                                                   float(grad_scale), _lib.current_stream_ptr()), "ay2_sgd_ema_step_groups")
    torch._C._increment_version([t for t in (param, mom, ema) if t is not None])  # takes an ITERABLE of tensors


def _dgrad_specs(weight: torch.Tensor, stride: int, pad: int):
    """[(sub-weight OIHW for the dgrad conv (out = cin, in = cout), kh, kw, pad_h, pad_w, out_sub)]."""
    w = weight.detach().float()
    cout, cin, kh, kw = w.shape
    if stride == 1:
        assert kh == kw
        return [(w.permute(1, 0, 2, 3).flip(2, 3).contiguous(), kh, kw, kh - 1 - pad, kw - 1 - pad, None)]
    assert stride == 2
    specs = []
    for py in range(2):
        khs = sorted([k for k in range(kh) if (k - pad - py) % 2 == 0], key=lambda k: (py + pad - k) // 2)
        offs_y = [(py + pad - k) // 2 for k in khs]
        for px in range(2):
            kws = sorted([k for k in range(kw) if (k - pad - px) % 2 == 0], key=lambda k: (px + pad - k) // 2)
            offs_x = [(px + pad - k) // 2 for k in kws]
            assert offs_y == list(range(offs_y[0], offs_y[0] + len(khs))) and offs_x == list(range(offs_x[0], offs_x[0] + len(kws)))
            # taps ordered by input offset == kernel index descending in steps of 2; slices + flip only (no index tensors:
            # this runs inside CUDA-graph capture, where a host->device index copy is illegal)
            assert khs == list(range(khs[0], khs[-1] - 1, -2)) and kws == list(range(kws[0], kws[-1] - 1, -2))
            sub = w[:, :, khs[-1]:khs[0] + 1:2, kws[-1]:kws[0] + 1:2].flip(2, 3)  # (cout, cin, len(khs), len(kws))
            specs.append((sub.permute(1, 0, 2, 3).contiguous(), len(khs), len(kws), -offs_y[0], -offs_x[0], (py, px)))
    return specs


def pixel_pair_weight(w: torch.Tensor) -> torch.Tensor:
    """OIHW weight of a 3x3 / stride-2 / pad-1 conv over 32 channels -> the (cout, 64, 3, 2) weight of the same conv read
    over PAIRS of pixels ([B, H, W/2, 64]; stride 2 over rows, 1 over pairs; one pair of zero padding on the left only): pair
    ox - 1 carries input column 2 ox - 1 in its upper 32 channels (kw = 0), pair ox columns 2 ox and 2 ox + 1 (kw = 1, 2)."""
    cout, cin, kh, kw = w.shape
    assert (cin, kh, kw) == (32, 3, 3), w.shape
    w2 = torch.zeros((cout, 64, 3, 2), dtype=w.dtype, device=w.device)
    w2[:, 32:, :, 0] = w[:, :, :, 0]
    w2[:, :32, :, 1] = w[:, :, :, 1]
    w2[:, 32:, :, 1] = w[:, :, :, 2]
    return w2


def conv_weight_layout(w: torch.Tensor) -> torch.Tensor:
    """OIHW fp32 -> the K-major operand layout [Cout_pad, KH*KW*Cin] in fp32 (pack_conv_weight without the BN fold and the
    bf16 cast): a pure index permutation + zero padding, so applying it to an index-valued tensor yields the gather table
    of `repack_weights`."""
    cout, cin, kh, kw = w.shape
    bn_tile = conv_block_n(cout)
    cout_pad = (cout + bn_tile - 1) // bn_tile * bn_tile
    out = torch.zeros((cout_pad, kh * kw * cin), dtype=torch.float32, device=w.device)
    out[:cout] = w.float().permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)
    return out


def dgrad_weight_layouts(weight: torch.Tensor, stride: int, pad: int, dz_channels: int) -> List[torch.Tensor]:
    """fp32 operand layouts of the dgrad launches, in the order make_dgrad_plans creates its plans (index permutations of
    `weight`, zero padded)."""
    out = []
    for wt, *_ in _dgrad_specs(weight, stride, pad):
        if wt.shape[1] < dz_channels:
            wt = torch.cat((wt, torch.zeros((wt.shape[0], dz_channels - wt.shape[1]) + tuple(wt.shape[2:]), device=wt.device)), 1)
        out.append(conv_weight_layout(wt))
    return out


def make_dgrad_weights(weight: torch.Tensor, stride: int, pad: int, dz_channels: int) -> List[torch.Tensor]:
    """Packed bf16 weights of the dgrad launches, in the order make_dgrad_plans creates its plans."""
    return [w.to(torch.bfloat16) for w in dgrad_weight_layouts(weight, stride, pad, dz_channels)]


def gather_index_of(layout_fn, shape, device) -> torch.Tensor:
    """int32 gather table of a layout function (OIHW fp32 -> operand layout): out.flat[i] = weight.flat[idx[i]], -1 = zero.
    Built by running the function on a tensor holding 1 + its own flat indices (exact in fp32 below 2^24 elements)."""
    n = 1
    for d in shape:
        n *= int(d)
    assert n < (1 << 24), "weight tensor too large for the fp32 index trick"
    probe = torch.arange(1, n + 1, dtype=torch.float32, device=device).view(*shape)
    return (layout_fn(probe).round().to(torch.int32) - 1).contiguous()


def repack_weights(table: torch.Tensor, nseg: int, total: int) -> None:
    """table: int64 (nseg, 4) CUDA tensor of (src fp32 ptr, dst bf16 ptr, int32 index ptr, first element); see include/ay2.h."""
    _lib.check(_lib.load().ay2_repack_weights(table.data_ptr(), nseg, total, _lib.current_stream_ptr()), "ay2_repack_weights")


def make_dgrad_plans(dz: ActView, dx: ActView, weight: torch.Tensor, stride: int, pad: int, accumulate: bool,
                     share: Optional[List["ConvPlan"]] = None) -> List[ConvPlan]:
    """Data gradient of `y = conv(x, weight, stride, pad)` as forward-conv launches of the same tcgen05 kernel:
       stride 1: dx = conv(dz, flip(W)^T, pad = k-1-p);
       stride 2: the four (row, column) parity sub-grids of dx are stride-1 convs of dz with the taps of matching
                 parity (ix = 2*ox + kw - p), written through a strided output view.
    weight: OIHW fp32 (the forward conv's). accumulate: dx += (fan-out), implemented with the residual input.
    share: plans of the other `accumulate` flavour of the same gradient, whose packed weight / bias buffers are reused."""
    cout, cin = weight.shape[0], weight.shape[1]
    assert dz.c >= cout and dx.c == cin, (dz.c, cout, dx.c, cin)
    res = dx if accumulate else None
    plans: List[ConvPlan] = []
    specs = _dgrad_specs(weight, stride, pad)
    packed = [pl.w for pl in share] if share is not None else make_dgrad_weights(weight, stride, pad, dz.c)
    for i, ((wt, kh, kw, ph, pw, sub), wp) in enumerate(zip(specs, packed)):
        bp = share[i].b if share is not None else torch.zeros(wp.shape[0], dtype=torch.float32, device=wp.device)
        plans.append(ConvPlan(dz, dx, wp, bp, kh, kw, 1, ph, ACT_NONE, residual=res, pad_w=pw, out_sub=sub))
    return plans
