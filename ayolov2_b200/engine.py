"""Execution engine: compiles a kindle-style YOLOModel into a fixed sequence of libay2 kernel launches.

What replaces what (reference call stack: scripts/utils/train_utils.py:403-472 validation_step ->
`model(imgs)` -> kindle per-layer PyTorch dispatch):
  * every `Conv` (conv + BN + SiLU)                 -> one fused tcgen05 implicit-GEMM launch (ay2_conv_plan_run)
  * `C3`: conv1 & conv2 share their input           -> ONE 1x1 conv with concatenated weights writing both halves of
                                                       the [.., 2c_] buffer conv3 reads (the torch.cat disappears);
                                                       bottleneck shortcut add fused into the 3x3 epilogue (in place)
  * `Concat` / the cat inside SPP(F)                -> addressing only: producers write channel slices
  * `UpSample`                                      -> one copy kernel straight into the concat slice
  * 6x6/s2 stem `Conv` and `Focus`                  -> space-to-depth kernel (also does uint8 -> /255 -> bf16) + 3x3 conv
  * `YOLOHead`                                      -> 1x1 conv per level + fused sigmoid/grid/anchor decode kernel
  * Tucker-2 decomposed convs (decomposition.py:363-424) -> three chained conv launches (ranks zero-padded to 16)
Activations are NHWC bf16; BN is folded with running statistics (eval mode).

The launch sequence is static, so a whole forward (+ NMS) is captured into one CUDA graph.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from . import ops
from .ops import ACT_NONE, ACT_SILU, ActView, ChainPlan, ConvPlan


def _act_code(m: nn.Module) -> int:
    a = getattr(m, "activation", None)
    if a is None or isinstance(a, nn.Identity):
        return ACT_NONE
    if isinstance(a, nn.SiLU):
        return ACT_SILU
    raise NotImplementedError(f"activation {type(a).__name__} is not implemented in the sm_100a conv epilogue (SiLU only)")


def _bn_tuple(bn: Optional[nn.Module]):
    if isinstance(bn, _BatchNorm):  # BatchNorm2d or SyncBatchNorm (train_model_builder.py:86-91 converts the model)
        return (bn.weight, bn.bias, bn.running_mean, bn.running_var), bn.eps
    return None, 1e-3


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


class Builder:
    """Accumulates launches (`steps`) and owns every buffer of one compiled graph."""

    def __init__(self, batch: int, device: torch.device, x3: bool = False) -> None:
        self.B = batch
        self.device = device
        # split-precision verification mode ("bf16x3", ops.ActView): same tcgen05 conv kernel, fp32-equivalent arithmetic.
        # Every producer writes ONE segment [hi | lo | hi]; `segments` remembers them per buffer so that a consumer can lay
        # its weights out as [w_hi | w_hi | w_lo] per input segment.
        self.x3 = x3
        self.segments: Dict[int, Dict[int, int]] = {}
        self.steps: List[Callable[[], None]] = []
        self.plans: List[ConvPlan] = []
        self.keep: List[Any] = []
        self.flops = 0.0
        self.act_bytes = 0.0  # algorithmic activation traffic (each conv reads its input once, writes its output once)

    def new_act(self, H: int, W: int, C_: int) -> ActView:
        v = ops.new_act(self.B, H, W, _round_up(C_, 8), device=self.device, x3=self.x3)
        v.buf.zero_()
        self.keep.append(v.buf)
        return v

    def mark_segment(self, v: ActView) -> None:
        if self.x3:
            self.segments.setdefault(v.buf.data_ptr(), {})[v.c0] = v.c

    def segments_of(self, v: ActView) -> List[Tuple[int, int]]:
        """[(offset inside the view, width)] of the producer segments that tile the split-precision view `v`."""
        table = self.segments.get(v.buf.data_ptr(), {})
        out, at = [], v.c0
        while at < v.c0 + v.c:
            assert at in table, f"split-precision view [{v.c0}, {v.c0 + v.c}) is not cut at producer segments {sorted(table.items())}"
            out.append((at - v.c0, table[at]))
            at += table[at]
        assert at == v.c0 + v.c, "split-precision view ends inside a producer segment"
        return out

    # ---------------------------------------------------------------------------------------------
    def conv2d(self, x: ActView, y: ActView, weight: torch.Tensor, bias: Optional[torch.Tensor], bn, eps: float,
               act: int, stride: int, pad: int, residual: Optional[ActView] = None, x2: Optional[ActView] = None) -> None:
        """weight OIHW fp32 (Cin may be smaller than x.c: zero-padded), output channels may be padded up to y.c.
        x2: the conv reads torch.cat([x, x2], 1) (two-source input, no concatenated tensor)."""
        w = weight.detach().float().to(self.device)
        cout, cin, kh, kw = w.shape
        xc = x.c + (x2.c if x2 is not None else 0)
        if cin < xc:
            assert x2 is None
            w = torch.cat((w, torch.zeros((cout, xc - cin, kh, kw), device=self.device)), 1)
        assert w.shape[1] == xc, (w.shape, xc)
        assert cout <= y.c
        if bn is not None:
            bn = tuple(t.detach().float().to(self.device) for t in bn)
        if bias is not None:
            bias = bias.detach().float().to(self.device)
        if cout < y.c:  # padded output channels produce act(0) = 0
            w = torch.cat((w, torch.zeros((y.c - cout,) + tuple(w.shape[1:]), device=self.device)), 0)
            if bn is not None:
                g, b_, mu, var = bn
                padn = y.c - cout
                bn = (torch.cat((g, torch.ones(padn, device=self.device))), torch.cat((b_, torch.zeros(padn, device=self.device))),
                      torch.cat((mu, torch.zeros(padn, device=self.device))), torch.cat((var, torch.ones(padn, device=self.device))))
            if bias is not None:
                bias = torch.cat((bias, torch.zeros(y.c - cout, device=self.device)))
        if (self.PIXEL_PAIRS and not self.x3 and (kh, kw, stride, pad) == (3, 3, 2, 1) and xc == 32 and x2 is None and residual is None
                and x.c0 == 0 and x.cstride == 32 and x.W % 2 == 0 and x.H % 2 == 0):
            # 3x3 / stride 2 over 32 channels (the layer after the stem): 64-byte operand rows, nine taps of K = 32. Read PAIRS
            # of pixels as 64-channel rows instead -- the same buffer viewed as [B, H, W/2, 64] -- with a 3x2-tap kernel:
            # pair ox-1 carries column 2ox-1 in its upper 32 channels (kw = 0), pair ox carries columns 2ox, 2ox+1 (kw = 1, 2).
            # Stride 2 over rows only: 128-byte rows, six taps (one weight block of zeros).
            wp, bp = ops.pack_conv_weight(ops.pixel_pair_weight(w), bias, bn, eps)
            xp = ActView(x.buf.view(x.B, x.H, x.W // 2, 64), 0, 64)
            plan = ConvPlan(xp, y, wp, bp, 3, 2, 2, 1, act, pad_w=1, stride_w=1)
            self.plans.append(plan)
            self.steps.append(plan.run)
            self.flops += 2.0 * self.B * y.H * y.W * cout * kh * kw * cin
            self.act_bytes += 2.0 * self.B * (x.H * x.W * cin + y.H * y.W * cout)
            return
        if self.x3:
            segs = self.segments_of(x)
            if x2 is not None:
                segs = segs + [(x.c + o, wd) for o, wd in self.segments_of(x2)]
            wp, bp = ops.pack_conv_weight_x3(w, bias, bn, eps, segs)
            self.mark_segment(y)
        else:
            wp, bp = ops.pack_conv_weight(w, bias, bn, eps)
        plan = ConvPlan(x, y, wp, bp, kh, kw, stride, pad, act, residual=residual, x2=x2)
        self.plans.append(plan)
        self.steps.append(plan.run)
        self.flops += 2.0 * self.B * y.H * y.W * cout * kh * kw * cin
        self.act_bytes += 2.0 * self.B * (x.H * x.W * cin + y.H * y.W * cout)

    # ---------------------------------------------------------------------------------------------
    # fused chains (csrc/conv_chain.cu): 1x1 -> 3x3 (-> 1x1) in one launch
    FUSE_CHAINS = True      # class-level switch (tests / A-B measurements)
    PIXEL_PAIRS = os.environ.get("AY2_PIXEL_PAIRS", "1") != "0"  # 3x3/s2 over 32 channels as a 3x2-tap conv over pixel pairs
    # Bottlenecks wider than this stay two launches: measured on B200 (r01, bs 64): c = 32 @160x160 fused 0.165 ms vs
    # 0.240 ms, c = 64 @80x80 0.097 vs 0.084 ms, c = 128 @40x40 0.105 vs 0.066 ms (the serialised stages of the fused
    # kernel lose to two pipelined launches once the 3x3 is tensor-bound). Tucker chains are always fused when supported.
    FUSE_BOTTLENECK_MAX_C = int(os.environ.get("AY2_FUSE_BOTTLENECK_MAX_C", "32"))
    CHAIN_MIN_TILE_EFF = 0.65  # 16x16 output tiles: below this coverage (20x20 maps: 0.39) separate launches win

    @staticmethod
    def _tile_eff(H: int, W: int) -> float:
        return (H * W) / float(((H + 15) // 16 * 16) * ((W + 15) // 16 * 16))

    def _chain_ok(self, x: ActView, c1: int, c2: int, c3: int, k: int, s: int, p: int) -> bool:
        if self.x3 or not self.FUSE_CHAINS or self._tile_eff(x.H, x.W) < self.CHAIN_MIN_TILE_EFF or x.c % 16:
            return False
        return ops.chain_supported(ops.chain_desc(x, c1, c2, c3, 0, 0, 0, 16, 0, k=k, stride=s, pad=p))

    def chain(self, x: ActView, y: ActView, links, acts, residual: Optional[ActView], flops: float, act_elems: float) -> None:
        """links: [(OIHW weight, bias, bn)] for the 2 or 3 convolutions; eps fixed by the caller's fold."""
        plan = ChainPlan(x, y, links, acts, residual=residual)
        self.plans.append(plan)
        self.steps.append(plan.run)
        self.flops += flops
        self.act_bytes += 2.0 * act_elems

    @staticmethod
    def _conv_geom(c: nn.Conv2d) -> Tuple[int, int]:
        assert c.groups == 1 and c.dilation == (1, 1), "grouped / dilated convolutions are not on the YOLOv5 hot path"
        assert c.stride[0] == c.stride[1] and c.padding[0] == c.padding[1] and c.kernel_size[0] == c.kernel_size[1]
        return c.stride[0], c.padding[0]

    def kindle_conv(self, m: nn.Module, x: ActView, y: Optional[ActView] = None, residual: Optional[ActView] = None) -> ActView:
        """kindle Conv (conv -> BN -> act); `m.conv` may be a Tucker-2 nn.Sequential of three Conv2d."""
        bn, eps = _bn_tuple(getattr(m, "batch_norm", None))
        act = _act_code(m)
        if isinstance(m.conv, nn.Sequential):  # decomposition.py:363-424: 1x1 (Cin->R1) -> kxk (R1->R0) -> 1x1 (R0->Cout, +bias)
            first, core, last = m.conv[0], m.conv[1], m.conv[2]
            s, p = self._conv_geom(core)
            k = core.kernel_size[0]
            oh, ow = (x.H + 2 * p - k) // s + 1, (x.W + 2 * p - k) // s + 1
            r1p, r0p = _round_up(first.out_channels, 16), _round_up(core.out_channels, 16)
            cout = last.out_channels
            if (first.bias is None and core.bias is None and cout % 16 == 0 and (y is None or y.c == cout)
                    and self._chain_ok(x, r1p, r0p, cout, k, s, p)):
                # ONE launch: the rank-R1 / rank-R0 intermediates stay in shared memory / TMEM
                if y is None:
                    y = self.new_act(oh, ow, cout)
                dev = self.device
                bn_d = None if bn is None else tuple(t.detach().float().to(dev) for t in bn)
                links = [ops.pack_chain_weight(first.weight.to(dev), None, cin_pad=x.c, cout_pad=r1p),
                         ops.pack_chain_weight(core.weight.to(dev), None, cin_pad=r1p, cout_pad=r0p),
                         ops.pack_chain_weight(last.weight.to(dev), None if last.bias is None else last.bias.to(dev), bn_d, eps,
                                               cin_pad=r0p, cout_pad=cout)]
                npx = self.B * x.H * x.W
                fl = 2.0 * npx * (first.in_channels * first.out_channels + k * k * core.in_channels * core.out_channels
                                  + last.in_channels * cout)
                self.chain(x, y, links, (ACT_NONE, ACT_NONE, act), residual, fl, npx * (first.in_channels + cout))
                return y
            t1 = self.new_act(x.H, x.W, _round_up(first.out_channels, 16))
            self.conv2d(x, t1, first.weight, first.bias, None, eps, ACT_NONE, 1, 0)
            t2 = self.new_act(oh, ow, _round_up(core.out_channels, 16))
            self.conv2d(t1, t2, core.weight, core.bias, None, eps, ACT_NONE, s, p)
            if y is None:
                y = self.new_act(oh, ow, last.out_channels)
            self.conv2d(t2, y, last.weight, last.bias, bn, eps, act, 1, 0, residual)
            return y
        c = m.conv
        s, p = self._conv_geom(c)
        k = c.kernel_size[0]
        if y is None:
            y = self.new_act((x.H + 2 * p - k) // s + 1, (x.W + 2 * p - k) // s + 1, c.out_channels)
        self.conv2d(x, y, c.weight, c.bias, bn, eps, act, s, p, residual)
        return y

    # ---------------------------------------------------------------------------------------------
    def stem(self, m: nn.Module, img_getter: Callable[[], torch.Tensor], H: int, W: int, scale: float,
             y: Optional[ActView]) -> ActView:
        """First layer fed by the NCHW image: 6x6/s2/p2 Conv or Focus(k, s=1) -> space-to-depth + k'xk' s1 conv."""
        name = type(m).__name__
        c = m.conv
        assert isinstance(c, nn.Conv2d), "a decomposed stem is not supported"
        # Space-to-depth output, horizontally padded: logical pixel x lives at physical column x + 1 and the row has
        # 8 spare columns, all zero. A k x k conv over the 16-channel pixels is then run as k x 1 taps over 4-pixel
        # WINDOWS (64 channels = one 128-byte TMA row) starting at physical column x: window = logical pixels x-1..x+2.
        H2, W2 = H // 2, W // 2
        Wp = W2 + 8
        s2d = ActView(torch.zeros((self.B, H2, Wp, 48 if self.x3 else 16), dtype=torch.bfloat16, device=self.device), 0, 16, self.x3)
        self.keep.append(s2d.buf)
        self.s2d_view, self.s2d_scale = s2d, scale  # the fused letterbox kernel (data_loader.PackedBatch) writes it directly
        if self.x3:
            inv = 1.0 / scale
            divisor = float(round(inv)) if abs(inv - round(inv)) < 1e-6 * inv else inv  # 1/255 -> exactly 255
            self.steps.append(lambda: ops.space_to_depth_x3(img_getter(), s2d, divisor, x_offset=1))
        else:
            self.steps.append(lambda: ops.space_to_depth(img_getter(), s2d, scale, x_offset=1))
        self.s2d_step = self.steps[-1]
        w = c.weight.detach().float().to(self.device)
        if name == "Focus":
            s, p = self._conv_geom(c)
            assert s == 1 and w.shape[1] == 12
            # Focus channel = (dx*2+dy)*3 + c  ->  s2d channel = (dy*2+dx)*3 + c
            w2 = torch.zeros_like(w)
            for dy in range(2):
                for dx in range(2):
                    w2[:, (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3] = w[:, (dx * 2 + dy) * 3:(dx * 2 + dy) * 3 + 3]
            k2, p2 = c.kernel_size[0], p
        elif name == "Conv" and c.kernel_size == (6, 6) and c.stride == (2, 2) and c.padding == (2, 2) and w.shape[1] == 3:
            # W3[n, (dy*2+dx)*3 + c, a, b] = W6[n, c, 2a+dy, 2b+dx]
            w2 = torch.zeros((w.shape[0], 12, 3, 3), device=self.device)
            for dy in range(2):
                for dx in range(2):
                    w2[:, (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3] = w[:, :, dy::2, dx::2]
            k2, p2 = 3, 1
        else:
            raise NotImplementedError("first layer must be the 6x6/s2/p2 Conv stem or Focus (SURVEY.md §8a M1/M2)")
        assert k2 == 3 and p2 == 1, "packed stem handles 3x3/p1 over the space-to-depth grid"
        cout = w.shape[0]
        bn, eps = _bn_tuple(getattr(m, "batch_norm", None))
        if bn is not None:
            bn = tuple(t.detach().float().to(self.device) for t in bn)
        if y is None:
            y = self.new_act(H2, W2, cout)
        bias = c.bias.detach().float().to(self.device) if c.bias is not None else None
        pair = (not self.x3 and y.c0 == 0 and y.cstride == cout and W2 % 2 == 0 and Wp % 2 == 0 and 2 * cout <= 256)
        if pair:
            # Two horizontally adjacent output pixels share one 4-pixel window (logical pixels 2x-1 .. 2x+2 are exactly
            # the union of their receptive fields): run the stem as M = pixel pairs, N = 2*cout. The window tile is
            # fetched once per pair, halving the L2->SM traffic of this L2-bound layer, and the 2*cout output channels
            # of a pair are the NHWC bytes of the two pixels, so the same buffer is simply viewed as [B,H,W/2,2*cout].
            ww = torch.zeros((2 * cout, 64, 3, 1), device=self.device)
            for px in range(2):
                for kw_ in range(3):
                    ww[px * cout:(px + 1) * cout, (kw_ + px) * 16:(kw_ + px) * 16 + 12, :, 0] = w2[:, :, :, kw_]
            bn2 = None if bn is None else tuple(torch.cat((t, t)) for t in bn)
            bias2 = None if bias is None else torch.cat((bias, bias))
            wp, bp = ops.pack_conv_weight(ww, bias2, bn2, eps)
            y2 = ActView(y.buf.view(self.B, H2, W2 // 2, 2 * cout), 0, 2 * cout)
            plan = ConvPlan(s2d, y2, wp, bp, 3, 1, 1, 1, _act_code(m), pad_w=0, window=(64, W2 // 2, 32, Wp // 2))
        else:
            # window weights: [cout, kh, (kwp in 0..3) * 16 + ch], kwp = 3 and ch >= 12 are zero
            ww = torch.zeros((cout, 64, 3, 1), device=self.device)  # OIHW with I = 64 window channels, kernel 3x1
            for kwp in range(3):
                ww[:, kwp * 16:kwp * 16 + 12, :, 0] = w2[:, :, :, kwp]
            if self.x3:  # the window's four pixels are four split-precision segments of 16 channels
                wp, bp = ops.pack_conv_weight_x3(ww, bias, bn, eps, [(0, 16), (16, 16), (32, 16), (48, 16)])
                self.mark_segment(y)
                plan = ConvPlan(s2d, y, wp, bp, 3, 1, 1, 1, _act_code(m), pad_w=0, window=(192, W2, 48, Wp))
            else:
                wp, bp = ops.pack_conv_weight(ww, bias, bn, eps)
                plan = ConvPlan(s2d, y, wp, bp, 3, 1, 1, 1, _act_code(m), pad_w=0, window=(64, W2, 16, Wp))
        self.plans.append(plan)
        self.steps.append(plan.run)
        self.flops += 2.0 * self.B * H2 * W2 * cout * 9 * 12  # == 36 taps x 3 channels of the 6x6 stem
        self.act_bytes += 2.0 * self.B * (H2 * W2 * 12 + H2 * W2 * cout)
        return y

    def bottleneck(self, m: nn.Module, y1: ActView) -> None:
        """In place on y1: y1 <- (y1 +) conv2_3x3(conv1_1x1(y1))."""
        t = self.kindle_conv(m.conv1, y1)
        self.kindle_conv(m.conv2, t, y=y1, residual=y1 if m.shortcut else None)

    def _bottleneck_fusable(self, m: nn.Module, x: ActView) -> bool:
        c1, c2 = m.conv1.conv, m.conv2.conv
        if not (isinstance(c1, nn.Conv2d) and isinstance(c2, nn.Conv2d)):
            return False
        if c1.kernel_size != (1, 1) or c1.stride != (1, 1) or c1.padding != (0, 0) or c1.groups != 1:
            return False
        if c2.kernel_size != (3, 3) or c2.stride != (1, 1) or c2.padding != (1, 1) or c2.groups != 1 or c2.dilation != (1, 1):
            return False
        if c1.in_channels != x.c or c2.out_channels != x.c or x.c > self.FUSE_BOTTLENECK_MAX_C:
            return False
        return self._chain_ok(x, c1.out_channels, c2.out_channels, 0, 3, 1, 1)

    def bottleneck_fused(self, m: nn.Module, x: ActView, y: ActView) -> None:
        """y <- (x +) conv2_3x3(conv1_1x1(x)) in ONE launch (y must be a different buffer: tiles read halos of x)."""
        dev = self.device
        links = []
        for cm in (m.conv1, m.conv2):
            bn, eps = _bn_tuple(getattr(cm, "batch_norm", None))
            bn = None if bn is None else tuple(t.detach().float().to(dev) for t in bn)
            bias = None if cm.conv.bias is None else cm.conv.bias.to(dev)
            links.append(ops.pack_chain_weight(cm.conv.weight.to(dev), bias, bn, eps))
        c_, ch = x.c, m.conv1.conv.out_channels
        npx = self.B * x.H * x.W
        self.chain(x, y, links, (_act_code(m.conv1), _act_code(m.conv2)), x if m.shortcut else None,
                   2.0 * npx * (c_ * ch + 9 * ch * c_), npx * (c_ + ch + ch + c_))

    def bottleneck_seq(self, blocks, cur: ActView, tmp: Optional[ActView] = None) -> ActView:
        """Runs the bottlenecks of a C3 / BottleneckCSP starting from `cur`; fused ones ping-pong between `cur`'s slice
        and one temporary (they cannot run in place; `tmp` may be a slice the caller placed). Returns the view holding
        the result."""
        home = cur
        for b in blocks:
            if self._bottleneck_fusable(b, cur):
                if cur is home:
                    if tmp is None:
                        tmp = self.new_act(cur.H, cur.W, cur.c)
                    dst = tmp
                else:
                    dst = home
                self.bottleneck_fused(b, cur, dst)
                cur = dst
            else:
                self.bottleneck(b, cur)
        return cur

    def c3(self, m: nn.Module, x: ActView, y: Optional[ActView]) -> ActView:
        c_ = m.conv1.conv.out_channels if isinstance(m.conv1.conv, nn.Conv2d) else m.conv1.conv[-1].out_channels
        if self.x3:
            # split precision: a view must be whole producer segments, so the two branches live in their own buffers
            # (no merged conv1 || conv2 launch) and conv3 reads them as two sources
            y1 = self.kindle_conv(m.conv1, x)
            y2 = self.kindle_conv(m.conv2, x)
            cur = self.bottleneck_seq(m.bottleneck_c3, y1)
            c3 = m.conv3
            if not isinstance(c3.conv, nn.Conv2d):
                raise NotImplementedError("split-precision mode: a decomposed C3.conv3 is not supported")
            bn, eps = _bn_tuple(getattr(c3, "batch_norm", None))
            s_, p_ = self._conv_geom(c3.conv)
            if y is None:
                y = self.new_act(x.H, x.W, c3.conv.out_channels)
            self.conv2d(cur, y, c3.conv.weight, c3.conv.bias, bn, eps, _act_code(c3), s_, p_, x2=y2)
            return y
        fusable = (isinstance(m.conv1.conv, nn.Conv2d) and isinstance(m.conv2.conv, nn.Conv2d)
                   and type(m.conv1.activation) is type(m.conv2.activation)
                   and type(getattr(m.conv1, "batch_norm", None)) is type(getattr(m.conv2, "batch_norm", None)))
        blocks = list(m.bottleneck_c3)
        c3m = m.conv3
        if (fusable and blocks and c_ % 32 == 0 and c_ <= self.FUSE_BOTTLENECK_MAX_C and self.FUSE_CHAINS
                and isinstance(c3m.conv, nn.Conv2d) and c3m.conv.kernel_size == (1, 1)):
            buf3 = self.new_act(x.H, x.W, 3 * c_)
            if self._bottleneck_fusable(blocks[0], buf3.slice(2 * c_, c_)):
                return self._c3_three_slices(m, x, y, c_, blocks, buf3)
        cat = self.new_act(x.H, x.W, 2 * c_)
        if fusable:
            c1, c2 = m.conv1, m.conv2
            w = torch.cat((c1.conv.weight.detach().float(), c2.conv.weight.detach().float()), 0)
            bn1, eps = _bn_tuple(getattr(c1, "batch_norm", None))
            bn2, _ = _bn_tuple(getattr(c2, "batch_norm", None))
            bn = None if bn1 is None else tuple(torch.cat((a.detach().float(), b.detach().float())) for a, b in zip(bn1, bn2))
            bias = None
            if c1.conv.bias is not None:
                bias = torch.cat((c1.conv.bias.detach().float(), c2.conv.bias.detach().float()))
            self.conv2d(x, cat, w, bias, bn, eps, _act_code(c1), 1, 0)
        else:
            self.kindle_conv(m.conv1, x, y=cat.slice(0, c_))
            self.kindle_conv(m.conv2, x, y=cat.slice(c_, c_))
        y1 = cat.slice(0, c_)
        cur = self.bottleneck_seq(m.bottleneck_c3, y1)
        if cur is y1:
            return self.kindle_conv(m.conv3, cat, y=y)
        # odd number of fused bottlenecks: the result sits in the temporary -> conv3 reads [tmp | cat[c_:]] as two sources
        c3 = m.conv3
        if not isinstance(c3.conv, nn.Conv2d) or c_ % 16:
            ops_copy = cur  # rare: copy back and take the ordinary path
            self.steps.append(lambda: y1.tensor().copy_(ops_copy.tensor()))
            return self.kindle_conv(m.conv3, cat, y=y)
        bn, eps = _bn_tuple(getattr(c3, "batch_norm", None))
        s_, p_ = self._conv_geom(c3.conv)
        if y is None:
            y = self.new_act(x.H, x.W, c3.conv.out_channels)
        self.conv2d(cur, y, c3.conv.weight, c3.conv.bias, bn, eps, _act_code(c3), s_, p_, x2=cat.slice(c_, c_))
        return y

    def _c3_three_slices(self, m: nn.Module, x: ActView, y: Optional[ActView], c_: int, blocks, buf: ActView) -> ActView:
        """C3 whose bottlenecks run as fused chain launches (which cannot work in place): ONE buffer of 3 c_ channels
        [R | Y2 | Y1]. The merged conv1 || conv2 launch writes [Y2 | Y1] (its filters in that order), the chain kernels
        ping-pong between Y1 and R, and conv3 reads ONE contiguous 2 c_-channel view -- [R | Y2] as is, or [Y2 | Y1] with
        the two halves of its input channels swapped -- instead of two sources with half-width (64-byte) operand rows,
        which ran at the TMA row rate (117 us vs 70 us for the same shape at 160 x 160)."""
        R, Y2, Y1 = buf.slice(0, c_), buf.slice(c_, c_), buf.slice(2 * c_, c_)
        c1, c2, c3 = m.conv1, m.conv2, m.conv3
        w = torch.cat((c2.conv.weight.detach().float(), c1.conv.weight.detach().float()), 0)
        bn1, eps = _bn_tuple(getattr(c1, "batch_norm", None))
        bn2, _ = _bn_tuple(getattr(c2, "batch_norm", None))
        bn = None if bn1 is None else tuple(torch.cat((b.detach().float(), a.detach().float())) for a, b in zip(bn1, bn2))
        bias = None
        if c1.conv.bias is not None:
            bias = torch.cat((c2.conv.bias.detach().float(), c1.conv.bias.detach().float()))
        self.conv2d(x, buf.slice(c_, 2 * c_), w, bias, bn, eps, _act_code(c1), 1, 0)
        cur = self.bottleneck_seq(blocks, Y1, tmp=R)
        bn3, eps3 = _bn_tuple(getattr(c3, "batch_norm", None))
        s_, p_ = self._conv_geom(c3.conv)
        if y is None:
            y = self.new_act(x.H, x.W, c3.conv.out_channels)
        w3 = c3.conv.weight.detach().float()
        if cur is R:
            src = buf.slice(0, 2 * c_)                                   # [R | Y2]: conv3's own channel order
        else:
            src = buf.slice(c_, 2 * c_)                                  # [Y2 | Y1]: swap the halves of the input channels
            w3 = torch.cat((w3[:, c_:], w3[:, :c_]), 1)
        self.conv2d(src, y, w3, c3.conv.bias, bn3, eps3, _act_code(c3), s_, p_)
        return y

    def bottleneck_csp(self, m: nn.Module, x: ActView, y: Optional[ActView]) -> ActView:
        c_ = m.conv1.conv.out_channels
        cat = self.new_act(x.H, x.W, 2 * c_)
        t = self.kindle_conv(m.conv1, x)
        t = self.bottleneck_seq(m.bottleneck_csp, t)
        # act(bn(cat[conv3(t), conv2(x)])) == cat[act(bn_a(conv3 t)), act(bn_b(conv2 x))]: fold each BN half
        bn, eps = _bn_tuple(m.batch_norm)
        act = _act_code(m)
        halves = [tuple(p[:c_] for p in bn), tuple(p[c_:] for p in bn)]
        self.conv2d(t, cat.slice(0, c_), m.conv3.weight, m.conv3.bias, halves[0], eps, act, 1, 0)
        self.conv2d(x, cat.slice(c_, c_), m.conv2.weight, m.conv2.bias, halves[1], eps, act, 1, 0)
        return self.kindle_conv(m.conv4, cat, y=y)

    def spp(self, m: nn.Module, x: ActView, y: Optional[ActView]) -> ActView:
        c_ = m.conv1.conv.out_channels if isinstance(m.conv1.conv, nn.Conv2d) else m.conv1.conv[-1].out_channels
        if type(m).__name__ == "SPPF":
            k = m.pooling.kernel_size
            ks = (k, 2 * k - 1, 3 * k - 2)  # p(p(x)) == window 2k-1, p(p(p(x))) == 3k-2 (stride 1, -inf padding)
        else:
            ks = tuple(int(p.kernel_size) for p in m.pooling_modules)
            assert len(ks) == 3, "SPP with other than 3 pooling windows is not implemented"
        cat = self.new_act(x.H, x.W, 4 * c_)
        self.kindle_conv(m.conv1, x, y=cat.slice(0, c_))
        s0, s1, s2, s3 = (cat.slice(i * c_, c_) for i in range(4))
        if self.x3:
            for sl in (s1, s2, s3):
                self.mark_segment(sl)
            self.steps.append(lambda: ops.sppf_pool_x3(s0, s1, s2, s3, ks))
        else:
            self.steps.append(lambda: ops.sppf_pool(s0, s1, s2, s3, ks))
        return self.kindle_conv(m.conv2, cat, y=y)

    def upsample(self, m: nn.Module, x: ActView, y: Optional[ActView]) -> ActView:
        assert float(m.scale_factor) == 2.0 and m.mode == "nearest", "only nearest x2 UpSample is on the YOLOv5 path"
        if y is None:
            y = self.new_act(2 * x.H, 2 * x.W, x.c)
        if self.x3:  # a segment is copied whole: the destination has the source's segment structure
            for off, wd in self.segments_of(x):
                self.mark_segment(y.slice(off, wd))
        self.steps.append(lambda: ops.upsample2x(x, y))
        return y


class Engine:
    """A compiled eval-mode forward of one YOLOModel for a fixed input shape.

    run(img) -> (pred [B, sum(na*ny*nx), no] fp32, [raw_i (B, na, ny, nx, no) fp32]) ; all asynchronous.
    """

    def __init__(self, model: nn.Module, batch: int, height: int, width: int, in_dtype: torch.dtype = torch.float32,
                 scale: float = 1.0, want_raw: bool = True, device: Optional[torch.device] = None,
                 use_graph: bool = True, precision: str = "bf16") -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("ayolov2_b200.Engine needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.B, self.H, self.W = batch, height, width
        self.in_dtype, self.scale = in_dtype, scale
        self.input = torch.zeros((batch, 3, height, width), dtype=in_dtype, device=self.device) if use_graph else None
        self._img = self.input
        if precision not in ("bf16", "bf16x3"):
            raise ValueError(f"precision {precision!r}: 'bf16' (the product path) or 'bf16x3' (split-precision verification mode)")
        self.precision = precision
        b = Builder(batch, self.device, x3=precision == "bf16x3")
        self.b = b
        layers = list(model.model)
        nL = len(layers)

        def src_of(i: int) -> List[int]:
            frm = getattr(layers[i], "from_idx", -1)
            frm = list(frm) if isinstance(frm, (list, tuple)) else [frm]
            return [i - 1 if f == -1 else f for f in frm]

        # Concat planning: producers write straight into their slice of the concat buffer
        dest: Dict[int, Tuple[int, int]] = {}
        for j in range(nL):
            if type(layers[j]).__name__ == "Concat":
                assert getattr(layers[j], "dimension", 1) == 1
                off = 0
                for s in src_of(j):
                    if s in dest or type(layers[s]).__name__ == "Concat":
                        raise NotImplementedError("a tensor feeding two Concat layers / nested Concat needs a copy kernel")
                    dest[s] = (j, off)
                    off += self._out_channels(layers, s, src_of)
        cat_bufs: Dict[int, ActView] = {}
        outs: List[Optional[ActView]] = [None] * nL
        self.pred: Optional[torch.Tensor] = None
        self.raw: List[torch.Tensor] = []

        def out_view(i: int, H: int, W: int, C_: int) -> Optional[ActView]:
            if i not in dest:
                return None
            j, off = dest[i]
            if j not in cat_bufs:
                cat_bufs[j] = b.new_act(H, W, self._out_channels(layers, j, src_of))
            cv = cat_bufs[j]
            assert (cv.H, cv.W) == (H, W)
            return cv.slice(off, C_)

        for i, m in enumerate(layers):
            name = type(m).__name__
            srcs = src_of(i)
            if name in ("Conv", "Focus") and srcs[0] < 0:
                co = m.conv.out_channels
                outs[i] = b.stem(m, lambda: self._img, height, width, scale, out_view(i, height // 2, width // 2, co))
                continue
            xin = [outs[s] for s in srcs]
            if name == "Conv":
                c = m.conv if isinstance(m.conv, nn.Conv2d) else m.conv[1]
                s_, p_ = Builder._conv_geom(c)
                k_ = c.kernel_size[0]
                oh, ow = (xin[0].H + 2 * p_ - k_) // s_ + 1, (xin[0].W + 2 * p_ - k_) // s_ + 1
                co = m.conv.out_channels if isinstance(m.conv, nn.Conv2d) else m.conv[-1].out_channels
                outs[i] = b.kindle_conv(m, xin[0], y=out_view(i, oh, ow, co))
            elif name == "C3":
                co = m.conv3.conv.out_channels if isinstance(m.conv3.conv, nn.Conv2d) else m.conv3.conv[-1].out_channels
                outs[i] = b.c3(m, xin[0], out_view(i, xin[0].H, xin[0].W, co))
            elif name == "BottleneckCSP":
                outs[i] = b.bottleneck_csp(m, xin[0], out_view(i, xin[0].H, xin[0].W, m.conv4.conv.out_channels))
            elif name in ("SPP", "SPPF"):
                co = m.conv2.conv.out_channels if isinstance(m.conv2.conv, nn.Conv2d) else m.conv2.conv[-1].out_channels
                outs[i] = b.spp(m, xin[0], out_view(i, xin[0].H, xin[0].W, co))
            elif name == "Upsample":
                outs[i] = b.upsample(m, xin[0], out_view(i, 2 * xin[0].H, 2 * xin[0].W, xin[0].c))
            elif name == "Concat":
                outs[i] = cat_bufs[i]
            elif name == "YOLOHead":
                self._head(m, xin, want_raw)
            else:
                raise NotImplementedError(f"layer {i} ({name}) is not on the YOLOv5 detection hot path")
        self.outs = outs
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.use_graph = use_graph
        self.post_steps: List[Callable[[], None]] = []  # e.g. NMS appended by Detector

    @staticmethod
    def _out_channels(layers, i: int, src_of) -> int:
        m = layers[i]
        name = type(m).__name__
        if name in ("Conv", "Focus"):
            return m.conv.out_channels if isinstance(m.conv, nn.Conv2d) else m.conv[-1].out_channels
        if name == "C3":
            return m.conv3.conv.out_channels if isinstance(m.conv3.conv, nn.Conv2d) else m.conv3.conv[-1].out_channels
        if name == "BottleneckCSP":
            return m.conv4.conv.out_channels
        if name in ("SPP", "SPPF"):
            return m.conv2.conv.out_channels if isinstance(m.conv2.conv, nn.Conv2d) else m.conv2.conv[-1].out_channels
        if name == "Upsample":
            return Engine._out_channels(layers, src_of(i)[0], src_of)
        if name == "Concat":
            return sum(Engine._out_channels(layers, s, src_of) for s in src_of(i))
        raise NotImplementedError(name)

    def _head(self, m: nn.Module, xs: Sequence[ActView], want_raw: bool) -> None:
        b = self.b
        na, no = m.na, m.no
        total = sum(na * x.H * x.W for x in xs)
        self.pred = torch.zeros((self.B, total, no), dtype=torch.float32, device=self.device)
        off = 0
        self.na, self.no = na, no
        self.decode_steps: List[Callable[[], None]] = []
        self.head_logits: List[ActView] = []
        self.head_plans: List[ConvPlan] = []  # the detect convolutions (Detector arms their candidate epilogue)
        self.head_row_off: List[int] = []
        self.head_strides = [float(s) for s in m.stride]
        self.head_anchors_px = [m.anchor_grid[i].detach().float().reshape(-1, 2).tolist() for i in range(len(xs))]
        for i, (conv, x) in enumerate(zip(m.conv, xs)):
            logits = b.new_act(x.H, x.W, _round_up(na * no, 16))
            b.conv2d(x, logits, conv.weight, conv.bias, None, 1e-3, ACT_NONE, 1, 0)
            self.head_plans.append(b.plans[-1])
            self.head_row_off.append(off)
            raw = torch.zeros((self.B, na, x.H, x.W, no), dtype=torch.float32, device=self.device) if want_raw else None
            if raw is not None:
                self.raw.append(raw)
            anchors_px = m.anchor_grid[i].detach().float().reshape(-1).contiguous().to(self.device)
            b.keep.append(anchors_px)
            stride = float(m.stride[i])
            xyxy = bool(getattr(m, "out_xyxy", False))
            if b.x3 or xyxy:
                b.steps.append(lambda l=logits, a=anchors_px, s=stride, o=off, r=raw: ops.head_decode2(l, na, no, s, a, self.pred, o, r, xyxy=xyxy))
            else:
                b.steps.append(lambda l=logits, a=anchors_px, s=stride, o=off, r=raw: ops.head_decode(l, na, no, s, a, self.pred, o, r))
            self.decode_steps.append(b.steps[-1])
            self.head_logits.append(logits)
            off += na * x.H * x.W

    # ---------------------------------------------------------------------------------------------
    def _launch_all(self) -> None:
        for s in self.b.steps:
            s()
        for s in self.post_steps:
            s()

    @property
    def n_launches(self) -> int:
        return len(self.b.steps)

    def run(self, img: Optional[torch.Tensor] = None):
        """img: NCHW [B,3,H,W] of the engine's dtype on the engine's device, or None to use `self.input` as is."""
        if self.use_graph:
            if img is not None and img.data_ptr() != self.input.data_ptr():
                self.input.copy_(img, non_blocking=True)
            self._img = self.input
            if self.graph is None:
                self._launch_all()  # warm-up (sets function attributes, loads modules) before capture
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch_all()
                self.graph = g
            self.graph.replay()
        else:
            self._img = self.input if img is None else img.contiguous()
            self._launch_all()
        return self.pred, self.raw


# -------------------------------------------------------------------------------------------------
# drop-in glue used by the kindle-compatible modules
# -------------------------------------------------------------------------------------------------
def _weights_signature(model: nn.Module) -> int:
    """Hash of (storage address, version counter) of every parameter and buffer: any in-place update, reallocation or
    re-binding of a tensor changes it (a sum of the fields would let two changes cancel)."""
    return hash(tuple((t.data_ptr(), t._version) for t in list(model.parameters()) + list(model.buffers())))


def set_precision(model: nn.Module, precision: str = "bf16") -> nn.Module:
    """Arithmetic of `model(x)` in eval mode: "bf16" (product path: bf16 storage, fp32 accumulation) or "bf16x3" (split-
    precision verification mode: every activation and weight carried as hi + lo bf16 pairs through the same tcgen05
    kernels, fp32-equivalent results at ~4x the cost; tests/test_precise_gpu.py)."""
    if precision not in ("bf16", "bf16x3"):
        raise ValueError(precision)
    model.__dict__["_ay2_precision"] = precision
    model.__dict__.get("_engine_cache", {}).clear()
    return model


def forward_model(model: nn.Module, x: torch.Tensor):
    """YOLOModel.forward for CUDA inputs (train.py / val.py call `model(imgs)`)."""
    if model.training:
        from .train_engine import forward_train

        return forward_train(model, x)
    if x.dim() != 4:
        raise ValueError("expected NCHW input")
    in_dtype = x.dtype
    scale = 1.0
    if x.dtype in (torch.float16, torch.bfloat16, torch.float64):
        x = x.float()
        in_dtype = torch.float32
    B, C_, H, W = x.shape
    precision = model.__dict__.get("_ay2_precision", "bf16")
    key = (B, H, W, in_dtype, x.device.index, precision)
    cache = model.__dict__.setdefault("_engine_cache", {})
    sig = _weights_signature(model)
    ent = cache.get(key)
    if ent is None or ent[1] != sig:
        cache.clear()
        ent = (Engine(model, B, H, W, in_dtype=in_dtype, scale=scale, device=x.device, use_graph=False, precision=precision), sig)
        cache[key] = ent
    eng = ent[0]
    pred, raw = eng.run(x)
    pred, raw = pred.clone(), [r.clone() for r in raw]
    if model.__dict__.get("_export", False):
        return (pred,)
    return pred, raw


def run_single_module(module: nn.Module, x):
    """`forward` of a bare kindle module on CUDA NCHW tensors: builds a one-layer plan (not cached; for the
    reference's module-level uses such as profiling — the model-level path is forward_model)."""
    name = type(module).__name__
    xs = list(x) if isinstance(x, (list, tuple)) else [x]
    if not all(t.is_cuda for t in xs):
        raise RuntimeError(f"ayolov2_b200 {name} runs on CUDA tensors only; there is no CPU fallback")
    if module.training and any(isinstance(mm, _BatchNorm) for mm in module.modules()):
        raise NotImplementedError("training-mode module forward is not built yet; call .eval()")
    dev = xs[0].device
    B = xs[0].shape[0]
    b = Builder(B, dev)
    views = []
    for t in xs:
        cpad = _round_up(t.shape[1], 16)
        buf = torch.zeros((B, t.shape[2], t.shape[3], cpad), dtype=torch.bfloat16, device=dev)
        buf[..., :t.shape[1]] = t.permute(0, 2, 3, 1)
        views.append(ActView(buf, 0, cpad))
    if name == "YOLOHead":
        class _E:  # minimal host for Engine._head
            pass
        e = Engine.__new__(Engine)
        e.b, e.B, e.device, e.raw, e.pred = b, B, dev, [], None
        e._head(module, views, True)
        for s in b.steps:
            s()
        return (e.pred, e.raw)
    if name == "Conv":
        y = b.kindle_conv(module, views[0])
    elif name == "Focus":
        raise NotImplementedError("Focus is compiled as part of a model (it consumes the NCHW image)")
    elif name == "Bottleneck":
        y = views[0]
        b.bottleneck(module, y)
    elif name == "C3":
        y = b.c3(module, views[0], None)
    elif name == "BottleneckCSP":
        y = b.bottleneck_csp(module, views[0], None)
    elif name in ("SPP", "SPPF"):
        y = b.spp(module, views[0], None)
    else:
        raise NotImplementedError(name)
    for s in b.steps:
        s()
    return y.tensor().permute(0, 3, 1, 2).float().contiguous()
