"""Training-mode execution: forward with batch-statistics BatchNorm and the full analytic backward of a
kindle-style YOLOModel, on the libay2 kernels.

Replaces, for `model.train()`, what scripts/train/yolo_trainer.py:322-329 runs through PyTorch/cuDNN:
`pred = model(imgs)` (per-layer conv -> BatchNorm2d(batch stats) -> SiLU) and the autograd backward that
`scaler.scale(loss).backward()` triggers. Per kindle Conv:

  forward   z  = conv(x, W)                        ay2_conv_plan_run (tcgen05 implicit GEMM, raw bf16 output)
            mu, 1/sigma, running stats             ay2_bn_stats + ay2_bn_finalize
            y  = SiLU(gamma*(z-mu)/sigma + beta)   ay2_bn_act_fwd (+ shortcut add)
  backward  dz, dgamma, dbeta                      ay2_bn_act_bwd
            dW += dz^T (*) x                       ay2_conv_wgrad (tcgen05, MN-major operands, split-K)
            dx += dz (*) W^T                       the forward kernel again with flipped/transposed weights
                                                   (stride 2: four parity sub-grids), accumulating via its residual input
Concat / C3 cat / SPP cat are channel slices of one buffer in both directions; UpSample and the max pools have
their own backward kernels; the YOLOHead is a 1x1 conv with bias whose gradient arrives as (bs, na, ny, nx, no).

`TrainFunction` exposes this as ONE torch.autograd.Function over all parameters, so `loss.backward()`, GradScaler,
DDP's gradient hooks and torch optimizers keep working unchanged (the drop-in contract of SURVEY.md §8b).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from . import ops
from .engine import _act_code, _round_up
from .ops import ACT_NONE, ActView, ConvPlan


class _ConvRec:
    """One convolution (+ optional BN/act) of the network with everything its backward needs."""

    def __init__(self) -> None:
        self.conv: nn.Conv2d = None  # type: ignore
        self.bn: Optional[nn.BatchNorm2d] = None
        self.act = ACT_NONE
        self.x: ActView = None  # type: ignore
        self.z: ActView = None  # type: ignore   raw conv output
        self.y: ActView = None  # type: ignore   bn/act output (== z when there is no BN)
        self.residual: Optional[ActView] = None
        self.stem = False


class _GradSite:
    """One write into a channel slice of a gradient buffer during the backward pass. The launch order of a backward is
    static, so the first (eager) backward decides once whether this site is the FIRST writer of its slice in a step --
    then it overwrites (no zero-fill of the buffer, no read of the old value) -- or a later one that accumulates."""

    def __init__(self, eng: "TrainEngine", gv: ActView):
        self.eng, self.key, self.lo, self.hi = eng, id(gv.buf), gv.c0, gv.c0 + gv.c
        self.acc: Optional[bool] = None

    def accumulate(self) -> bool:
        if self.acc is None:
            self.acc = self.eng._claim(self.key, self.lo, self.hi)
        return self.acc


class TrainEngine:
    def __init__(self, model: nn.Module, batch: int, height: int, width: int, in_dtype: torch.dtype = torch.float32,
                 scale: float = 1.0, device: Optional[torch.device] = None, use_graph: bool = True) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("ayolov2_b200.TrainEngine needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.B, self.H, self.W = batch, height, width
        self.scale = scale
        self.fwd: List[Callable[[], None]] = []
        self.bwd: List[Callable[[], None]] = []   # appended in forward order, executed reversed
        self.refresh: List[Callable[[], None]] = []  # weight re-packing (parameters change every optimizer step)
        self.keep: List[Any] = []
        self.grad_of: Dict[int, torch.Tensor] = {}   # id(activation buffer) -> gradient buffer
        self.overwritten: set = set()                 # activation buffers whose gradient needs no zero-fill per step
        self._cover: Dict[int, List[Tuple[int, int]]] = {}  # gradient buffer -> channel intervals written so far (first backward)
        self._needs_zero: set = set()                 # gradient buffers with partially overlapping writers: zero-filled per step
        self._gbuf: Dict[int, torch.Tensor] = {}      # id(gradient buffer) -> gradient buffer
        self._repack: List[Tuple[nn.Parameter, torch.Tensor, torch.Tensor]] = []  # (fp32 parameter, packed bf16 operand, gather index)
        self._repack_table: Optional[torch.Tensor] = None
        self._repack_total, self._repack_ptrs = 0, None
        self._bn_counters: List[torch.Tensor] = []    # num_batches_tracked of every BatchNorm the forward passes through
        # per-layer BatchNorm sums (statistics in the forward, the two backward sums later) are slices of ONE double buffer,
        # cleared by one fill at the start of each pass instead of one fill / memset per layer
        self._sums_flat = torch.zeros(1 << 18, dtype=torch.float64, device=self.device)
        self._sums_used = 0
        self._img: Optional[torch.Tensor] = None
        self.head_out: List[torch.Tensor] = []
        self.head_gin: List[torch.Tensor] = []
        self.flops_fwd = 0.0
        # The launch sequence of a step is static (fixed shapes, static buffers, parameters updated in place), so the
        # forward and the backward are each captured into a CUDA graph after one eager warm-up call: ~2000 small
        # launches (weight re-packing, per-layer kernels, gradient scatter) become two graph replays.
        self.use_graph = use_graph
        self._gstate: Dict[str, int] = {}
        self._graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self.static_in = torch.zeros((batch, 3, height, width), dtype=in_dtype, device=self.device)
        self._build()
        # parameter gradients: views (in model.parameters() order) of ONE flat fp32 buffer, so that a step zeroes, copies,
        # all-reduces and applies them with a handful of launches instead of one per parameter
        params = list(model.parameters())
        self.pg_offsets: List[int] = []
        off = 0
        for p in params:
            self.pg_offsets.append(off)
            off += _round_up(p.numel(), 4)  # 16-byte aligned slices
        self.pg_flat = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.pg: Dict[int, torch.Tensor] = {id(p): self.pg_flat[o:o + p.numel()].view(p.shape) for p, o in zip(params, self.pg_offsets)}
        self.last_grad_flat: Optional[torch.Tensor] = None  # the flat gradient of the latest backward (same layout)

    # ------------------------------------------------------------------------------------------------ buffers
    def new_act(self, H: int, W: int, C_: int) -> ActView:
        v = ops.new_act(self.B, H, W, _round_up(C_, 8), device=self.device)
        v.buf.zero_()
        self.keep.append(v.buf)
        return v

    def g(self, v: ActView) -> ActView:
        """Gradient view mirroring an activation view."""
        gb = self.grad_of.get(id(v.buf))
        if gb is None:
            gb = torch.zeros_like(v.buf)
            self.grad_of[id(v.buf)] = gb
            self._gbuf[id(gb)] = gb
        return ActView(gb, v.c0, v.c)

    def _claim(self, key: int, lo: int, hi: int) -> bool:
        """Resolution of a _GradSite at the first backward: False = first writer of [lo, hi) (overwrite), True = accumulate."""
        iv = self._cover.setdefault(key, [])
        hit = sorted((a, b) for a, b in iv if a < hi and b > lo)
        if not hit:
            iv.append((lo, hi))
            return False
        pos = lo
        for a, b in hit:
            if a > pos:
                break
            pos = max(pos, b)
        if pos < hi:  # partly fresh, partly written: accumulate onto a per-step zero-fill (never the case in the YOLOv5 graphs)
            self._needs_zero.add(key)
            iv.append((lo, hi))
        return True

    # ------------------------------------------------------------------------------------------------ weight re-packing
    def _repack_add(self, param: nn.Parameter, dst: torch.Tensor, layout_fn) -> None:
        """Register a packed bf16 operand `dst` = layout_fn(param) (an index permutation with zero padding): all of them are
        refreshed from the current parameter values by ONE gather launch at the start of every forward."""
        idx = ops.gather_index_of(layout_fn, tuple(param.shape), self.device)
        assert idx.numel() == dst.numel() and dst.is_contiguous() and dst.numel() % 8 == 0, (idx.shape, dst.shape)
        self._repack.append((param, dst, idx))

    def _repack_run(self) -> None:
        if self._repack:
            ops.repack_weights(self._repack_table, len(self._repack), self._repack_total)

    def _repack_sync_table(self) -> None:
        """(Re)build the device segment table when a parameter's storage moved (first call; TrainStep re-points the
        parameters into its flat buffer; `.to()` / `load_state_dict` with assign). Outside the CUDA graphs: the captured
        gather reads the table at run time."""
        ptrs = [p.data_ptr() for p, _, _ in self._repack]
        if ptrs == self._repack_ptrs:
            return
        rows, begin = [], 0
        for p, dst, idx in self._repack:
            assert p.dtype == torch.float32 and p.is_contiguous()
            rows.append((p.data_ptr(), dst.data_ptr(), idx.data_ptr(), begin))
            begin += dst.numel()
        host = torch.tensor(rows, dtype=torch.int64)
        if self._repack_table is None:
            self._repack_table = host.to(self.device)
        else:
            self._repack_table.copy_(host)
        self._repack_total, self._repack_ptrs = begin, ptrs

    def _sums(self, n: int) -> torch.Tensor:
        n = _round_up(n, 2)
        assert self._sums_used + n <= self._sums_flat.numel(), "BatchNorm sum buffer too small for this model"
        v = self._sums_flat[self._sums_used:self._sums_used + n]
        self._sums_used += n
        return v

    def _padd(self, p: nn.Parameter, t: torch.Tensor) -> None:
        self.pg[id(p)].add_(t.reshape(p.shape))

    # ------------------------------------------------------------------------------------------------ conv + bn + act
    def conv_bn_act(self, conv: nn.Conv2d, bn: Optional[nn.Module], act: int, x: ActView, y: Optional[ActView] = None,
                    residual: Optional[ActView] = None, need_dx: bool = True) -> ActView:
        assert conv.groups == 1 and conv.dilation == (1, 1) and conv.kernel_size[0] == conv.kernel_size[1]
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        cout, cin = conv.out_channels, conv.in_channels
        assert x.c == cin, (x.c, cin)
        oh, ow = (x.H + 2 * p - k) // s + 1, (x.W + 2 * p - k) // s + 1
        has_bn = isinstance(bn, _BatchNorm)
        sync = isinstance(bn, nn.SyncBatchNorm)  # statistics and their backward sums over the cross-rank batch
        if y is None:
            y = self.new_act(oh, ow, cout)
        z = self.new_act(oh, ow, cout) if has_bn else y
        dev = self.device
        bn_tile = ops.conv_block_n(cout)
        cout_pad = _round_up(cout, bn_tile)
        wp = torch.zeros((cout_pad, k * k * cin), dtype=torch.bfloat16, device=dev)
        bp = torch.zeros(cout_pad, dtype=torch.float32, device=dev)
        plan = ConvPlan(x, z, wp, bp, k, k, s, p, ACT_NONE if has_bn else act)
        self.keep += [wp, bp, plan]
        self.flops_fwd += plan.flops

        self._repack_add(conv.weight, wp, ops.conv_weight_layout)
        if conv.bias is not None:
            self.refresh.append(lambda: bp[:cout].copy_(conv.bias.detach()))
        self.fwd.append(plan.run)
        if has_bn:
            self._bn_counters.append(bn.num_batches_tracked)  # all incremented by one launch at the end of the forward
            mean = torch.empty(cout, device=dev)
            invstd = torch.empty(cout, device=dev)
            scratch = self._sums(2 * cout)
            self.keep += [mean, invstd]

            def f_bn() -> None:
                ops.bn_batch_stats(z, bn.eps, bn.momentum, bn.running_mean, bn.running_var, scratch, mean, invstd, sync=sync, zeroed=True)
                ops.bn_act_fwd(z, mean, invstd, bn.weight.detach().float(), bn.bias.detach().float(), act, y, residual)
            self.fwd.append(f_bn)
        else:
            assert residual is None and act == ACT_NONE, "conv without BN is only used by the head"
        # ---------------- backward
        gy, gz = self.g(y), (self.g(z) if has_bn else self.g(y))
        if has_bn:
            self.overwritten.add(id(z.buf))
        gx = self.g(x) if need_dx else None
        dw = torch.zeros((cout, k * k * cin), dtype=torch.float32, device=dev)
        dplans: List[ConvPlan] = []
        dplans_first: List[ConvPlan] = []
        dx_site = _GradSite(self, gx) if need_dx else None
        if need_dx:
            dplans = ops.make_dgrad_plans(gz, gx, conv.weight, s, p, accumulate=True)
            dplans_first = ops.make_dgrad_plans(gz, gx, conv.weight, s, p, accumulate=False, share=dplans)
            self.keep += dplans + dplans_first
            for i, pl in enumerate(dplans):
                self._repack_add(conv.weight, pl.w, lambda w, i=i: ops.dgrad_weight_layouts(w, s, p, gz.c)[i])
        gres = self.g(residual) if residual is not None else None
        res_site = _GradSite(self, gres) if gres is not None else None

        def b() -> None:
            if has_bn:
                if gres is not None:
                    ops.add_slices(gy, gres, accumulate=res_site.accumulate())  # shortcut branch: d(residual) (+)= dy
                # (d beta, d gamma go straight into the flat gradient from the apply pass unless SyncBatchNorm has to
                # all-reduce the sums first)
                if not ops.bn_act_bwd(gy, z, mean, invstd, bn.weight.detach().float(), bn.bias.detach().float(), act, scratch, gz,
                                      sync=sync, grad_beta=self.pg[id(bn.bias)], grad_gamma=self.pg[id(bn.weight)], zeroed=True):
                    self._padd(bn.bias, scratch[:cout].float())
                    self._padd(bn.weight, scratch[cout:].float())
            if k == 1:  # [cout][cin] IS the OIHW layout: accumulate straight into the flat gradient (zeroed once per step)
                ops.conv_wgrad(x, gz, self.pg[id(conv.weight)].view(cout, cin), 1, 1, s, p)
            else:
                dw.zero_()
                ops.conv_wgrad(x, gz, dw, k, k, s, p)
                self._padd(conv.weight, dw.view(cout, k, k, cin).permute(0, 3, 1, 2))
            if conv.bias is not None:
                bs = torch.zeros(_round_up(cout, 8), dtype=torch.float64, device=dev)
                ops.channel_sum(ActView(gz.buf, gz.c0, _round_up(cout, 8)), bs)
                self._padd(conv.bias, bs[:cout].float())
            if need_dx:
                for pl in (dplans if dx_site.accumulate() else dplans_first):
                    pl.run()
        self.bwd.append(b)
        return y

    def kindle_conv(self, m: nn.Module, x: ActView, y: Optional[ActView] = None, residual: Optional[ActView] = None) -> ActView:
        if isinstance(m.conv, nn.Sequential):
            # Tucker-2 chain (scripts/tensor_decomposition/decomposition.py:363-424): 1x1 -> kxk -> 1x1 plain convolutions,
            # BatchNorm + activation after the last one. Fine-tuning runs the links as separate launches (forward, data and
            # weight gradients each); the fused chain kernel is an eval-mode kernel.
            convs = list(m.conv)
            if not all(isinstance(c, nn.Conv2d) for c in convs) or any(c.out_channels % 8 for c in convs[:-1]):
                raise NotImplementedError("training a decomposed Conv needs nn.Conv2d links whose ranks are multiples of 8 "
                                          f"(got {[getattr(c, 'out_channels', type(c).__name__) for c in convs]})")
            cur = x
            for c in convs[:-1]:
                cur = self.conv_bn_act(c, None, ACT_NONE, cur)
            return self.conv_bn_act(convs[-1], getattr(m, "batch_norm", None), _act_code(m), cur, y, residual)
        return self.conv_bn_act(m.conv, getattr(m, "batch_norm", None), _act_code(m), x, y, residual)

    # ------------------------------------------------------------------------------------------------ stem
    def stem(self, m: nn.Module, y: Optional[ActView]) -> ActView:
        """6x6/s2/p2 Conv or Focus fed by the image: space-to-depth + packed 3x1-tap window conv (see engine.Builder.stem).
        Backward: weight gradient over the 16-channel space-to-depth image (3x3 taps), mapped back to the 6x6 / Focus
        layout on the host; no data gradient (the input is the image)."""
        name = type(m).__name__
        c = m.conv
        bn = getattr(m, "batch_norm", None)
        assert isinstance(c, nn.Conv2d) and isinstance(bn, _BatchNorm)
        sync = isinstance(bn, nn.SyncBatchNorm)
        H2, W2 = self.H // 2, self.W // 2
        Wp = W2 + 8
        dev = self.device
        s2d = ActView(torch.zeros((self.B, H2, Wp, 16), dtype=torch.bfloat16, device=dev), 0, 16)
        self.keep.append(s2d.buf)
        self.fwd.append(lambda: ops.space_to_depth(self._img, s2d, self.scale, x_offset=1))
        cout = c.out_channels
        is_focus = name == "Focus"
        if is_focus:
            assert c.kernel_size == (3, 3) and c.stride == (1, 1) and c.padding == (1, 1) and c.in_channels == 12
        else:
            assert c.kernel_size == (6, 6) and c.stride == (2, 2) and c.padding == (2, 2) and c.in_channels == 3

        def to_s2d_weight(w: torch.Tensor) -> torch.Tensor:  # -> (cout, 12, 3, 3) over s2d channels (dy*2+dx)*3+c
            w2 = torch.zeros((cout, 12, 3, 3), device=dev)
            for dy in range(2):
                for dx in range(2):
                    blk = slice((dy * 2 + dx) * 3, (dy * 2 + dx) * 3 + 3)
                    if is_focus:
                        w2[:, blk] = w[:, (dx * 2 + dy) * 3:(dx * 2 + dy) * 3 + 3]
                    else:
                        w2[:, blk] = w[:, :, dy::2, dx::2]
            return w2

        def from_s2d_grad(g2: torch.Tensor) -> torch.Tensor:  # inverse mapping for the gradient
            gw = torch.zeros_like(c.weight, dtype=torch.float32)
            for dy in range(2):
                for dx in range(2):
                    blk = slice((dy * 2 + dx) * 3, (dy * 2 + dx) * 3 + 3)
                    if is_focus:
                        gw[:, (dx * 2 + dy) * 3:(dx * 2 + dy) * 3 + 3] = g2[:, blk]
                    else:
                        gw[:, :, dy::2, dx::2] = g2[:, blk]
            return gw

        if y is None:
            y = self.new_act(H2, W2, cout)
        z = self.new_act(H2, W2, cout)
        # pair form (engine.Builder.stem): two horizontally adjacent output pixels share one 4-pixel window, M = pixel pairs,
        # N = 2 * cout -- the window tile is fetched once per pair and the MMA is twice as wide
        pair = z.c0 == 0 and z.cstride == cout and W2 % 2 == 0 and Wp % 2 == 0 and 2 * cout <= 256
        npx = 2 if pair else 1
        bn_tile = ops.conv_block_n(npx * cout)
        wp = torch.zeros((_round_up(npx * cout, bn_tile), 3 * 64), dtype=torch.bfloat16, device=dev)
        bp = torch.zeros(_round_up(npx * cout, bn_tile), dtype=torch.float32, device=dev)
        if pair:
            z2 = ActView(z.buf.view(self.B, H2, W2 // 2, 2 * cout), 0, 2 * cout)
            plan = ConvPlan(s2d, z2, wp, bp, 3, 1, 1, 1, ACT_NONE, pad_w=0, window=(64, W2 // 2, 32, Wp // 2))
        else:
            plan = ConvPlan(s2d, z, wp, bp, 3, 1, 1, 1, ACT_NONE, pad_w=0, window=(64, W2, 16, Wp))
        self.keep += [wp, bp, plan]
        self.flops_fwd += 2.0 * self.B * H2 * W2 * cout * 108

        def refresh() -> None:
            w2 = to_s2d_weight(c.weight.detach().float())
            ww = torch.zeros((npx * cout, 3, 64), device=dev)  # [pixel of the pair * cout + o][kh][window pixel * 16 + ch]
            for px in range(npx):
                for kwp in range(3):
                    ww[px * cout:(px + 1) * cout, :, (kwp + px) * 16:(kwp + px) * 16 + 12] = w2[:, :, :, kwp].permute(0, 2, 1)
            wp[:npx * cout].copy_(ww.reshape(npx * cout, -1))
        self.refresh.append(refresh)
        self.fwd.append(plan.run)
        mean, invstd = torch.empty(cout, device=dev), torch.empty(cout, device=dev)
        scratch = self._sums(2 * cout)
        act = _act_code(m)
        self._bn_counters.append(bn.num_batches_tracked)

        def f_bn() -> None:
            ops.bn_batch_stats(z, bn.eps, bn.momentum, bn.running_mean, bn.running_var, scratch, mean, invstd, sync=sync, zeroed=True)
            ops.bn_act_fwd(z, mean, invstd, bn.weight.detach().float(), bn.bias.detach().float(), act, y, None)
        self.fwd.append(f_bn)
        gy, gz = self.g(y), self.g(z)
        self.overwritten.add(id(z.buf))
        dw = torch.zeros((cout, 9 * 16), dtype=torch.float32, device=dev)
        x_ptr = s2d.ptr() + 2 * 16  # logical pixel 0 lives at physical column 1

        def b() -> None:
            if not ops.bn_act_bwd(gy, z, mean, invstd, bn.weight.detach().float(), bn.bias.detach().float(), act, scratch, gz,
                                  sync=sync, grad_beta=self.pg[id(bn.bias)], grad_gamma=self.pg[id(bn.weight)], zeroed=True):
                self._padd(bn.bias, scratch[:cout].float())
                self._padd(bn.weight, scratch[cout:].float())
            dw.zero_()
            ops.conv_wgrad(s2d, gz, dw, 3, 3, 1, 1, in_row_pixels=Wp, x_ptr=x_ptr, in_w=W2)
            g2 = dw.view(cout, 3, 3, 16)[..., :12].permute(0, 3, 1, 2)  # (cout, 12, 3, 3)
            self._padd(c.weight, from_s2d_grad(g2))
        self.bwd.append(b)
        return y

    # ------------------------------------------------------------------------------------------------ composite modules
    def bottleneck(self, m: nn.Module, x: ActView, y: Optional[ActView]) -> ActView:
        t = self.kindle_conv(m.conv1, x)
        return self.kindle_conv(m.conv2, t, y=y, residual=x if m.shortcut else None)

    def c3(self, m: nn.Module, x: ActView, y: Optional[ActView]) -> ActView:
        c_ = m.conv1.conv.out_channels
        cat = self.new_act(x.H, x.W, 2 * c_)
        n = len(m.bottleneck_c3)
        cur = self.kindle_conv(m.conv1, x, y=cat.slice(0, c_) if n == 0 else None)
        for i, b in enumerate(m.bottleneck_c3):
            cur = self.bottleneck(b, cur, cat.slice(0, c_) if i == n - 1 else None)
        self.kindle_conv(m.conv2, x, y=cat.slice(c_, c_))
        return self.kindle_conv(m.conv3, cat, y=y)

    def spp(self, m: nn.Module, x: ActView, y: Optional[ActView]) -> ActView:
        c_ = m.conv1.conv.out_channels
        cascade = type(m).__name__ == "SPPF"
        if cascade:
            k = m.pooling.kernel_size
            ks = (k, 2 * k - 1, 3 * k - 2)
        else:
            ks = tuple(int(p.kernel_size) for p in m.pooling_modules)
        cat = self.new_act(x.H, x.W, 4 * c_)
        sl = [cat.slice(i * c_, c_) for i in range(4)]
        self.kindle_conv(m.conv1, x, y=sl[0])
        self.fwd.append(lambda: ops.sppf_pool(sl[0], sl[1], sl[2], sl[3], ks))
        out = self.kindle_conv(m.conv2, cat, y=y)
        gs = [self.g(s) for s in sl]
        sites = [_GradSite(self, gs[2]), _GradSite(self, gs[1]), _GradSite(self, gs[0])] if cascade else \
            [_GradSite(self, gs[0]) for _ in range(3)]

        def b() -> None:  # runs after conv2's backward filled d(cat)
            if cascade:  # p3 = pool(p2), p2 = pool(p1), p1 = pool(x1): route through the cascade like the reference
                ops.maxpool_bwd(sl[2], gs[3], ks[0], gs[2], accumulate=sites[0].accumulate())
                ops.maxpool_bwd(sl[1], gs[2], ks[0], gs[1], accumulate=sites[1].accumulate())
                ops.maxpool_bwd(sl[0], gs[1], ks[0], gs[0], accumulate=sites[2].accumulate())
            else:
                for i in (1, 2, 3):
                    ops.maxpool_bwd(sl[0], gs[i], ks[i - 1], gs[0], accumulate=sites[i - 1].accumulate())
        # order: this must run BEFORE conv1's backward and AFTER conv2's: conv2's b() was appended last, so insert
        # the pool backward just before it in list order (lists are executed reversed)
        self.bwd.insert(len(self.bwd) - 1, b)
        return out

    def upsample(self, m: nn.Module, x: ActView, y: Optional[ActView]) -> ActView:
        assert float(m.scale_factor) == 2.0 and m.mode == "nearest"
        if y is None:
            y = self.new_act(2 * x.H, 2 * x.W, x.c)
        self.fwd.append(lambda: ops.upsample2x(x, y))
        gy, gx = self.g(y), self.g(x)
        site = _GradSite(self, gx)
        self.bwd.append(lambda: ops.upsample2x_bwd(gy, gx, accumulate=site.accumulate()))
        return y

    def head(self, m: nn.Module, xs: Sequence[ActView]) -> None:
        na, no = m.na, m.no
        for i, (conv, x) in enumerate(zip(m.conv, xs)):
            cpad = _round_up(na * no, 16)
            logits = self.new_act(x.H, x.W, cpad)
            out = torch.zeros((self.B, na, x.H, x.W, no), dtype=torch.float32, device=self.device)
            self.head_out.append(out)
            gl = self.g(logits)
            # The gradient arrives as (bs, na, ny, nx, no) fp32 and is converted to NHWC bf16 straight from the caller's
            # tensor, eagerly, before the backward graphs replay (`backward`): staging it in a static buffer first would
            # be a 1.1 GB read + write per step at bs 128. The conversion writes every channel of `gl`, padding included.
            self.head_gin.append(gl)
            self.overwritten.add(id(logits.buf))
            self._head_conv(conv, x, logits, na * no)
            self.fwd.append(lambda l=logits, o=out: ops.head_logits_to_train(l, na, no, o))

    def _head_conv(self, conv: nn.Conv2d, x: ActView, logits: ActView, nch: int) -> None:
        """1x1 conv with bias, Cout = na*no padded to the logits buffer width."""
        dev = self.device
        cin = conv.in_channels
        cpad = logits.c
        bn_tile = ops.conv_block_n(cpad)
        wp = torch.zeros((_round_up(cpad, bn_tile), cin), dtype=torch.bfloat16, device=dev)
        bp = torch.zeros(_round_up(cpad, bn_tile), dtype=torch.float32, device=dev)
        plan = ConvPlan(x, logits, wp, bp, 1, 1, 1, 0, ACT_NONE)
        self.keep += [wp, bp, plan]
        self.flops_fwd += 2.0 * self.B * x.H * x.W * nch * cin

        def refresh() -> None:
            wp[:nch].copy_(conv.weight.detach().reshape(nch, cin))
            bp[:nch].copy_(conv.bias.detach())
        self.refresh.append(refresh)
        self.fwd.append(plan.run)
        gl, gx = self.g(logits), self.g(x)
        dw = torch.zeros((cpad, cin), dtype=torch.float32, device=dev)
        wfull = torch.zeros((cpad, cin, 1, 1), device=dev)
        dplans = ops.make_dgrad_plans(gl, gx, wfull, 1, 0, accumulate=True)
        dplans_first = ops.make_dgrad_plans(gl, gx, wfull, 1, 0, accumulate=False, share=dplans)
        dx_site = _GradSite(self, gx)
        self.keep += dplans + dplans_first

        def refresh_d() -> None:
            wfull[:nch].copy_(conv.weight.detach())
            for pl, wnew in zip(dplans, ops.make_dgrad_weights(wfull, 1, 0, gl.c)):
                pl.w.copy_(wnew)
        self.refresh.append(refresh_d)

        def b() -> None:
            dw.zero_()
            ops.conv_wgrad(x, gl, dw, 1, 1, 1, 0)
            self._padd(conv.weight, dw[:nch].view(nch, cin, 1, 1))
            bs = torch.zeros(cpad, dtype=torch.float64, device=dev)
            ops.channel_sum(gl, bs)
            self._padd(conv.bias, bs[:nch].float())
            for pl in (dplans if dx_site.accumulate() else dplans_first):
                pl.run()
        self.bwd.append(b)

    # ------------------------------------------------------------------------------------------------ graph walk
    def _build(self) -> None:
        from .engine import Engine

        layers = list(self.model.model)
        nL = len(layers)

        def src_of(i: int) -> List[int]:
            frm = getattr(layers[i], "from_idx", -1)
            frm = list(frm) if isinstance(frm, (list, tuple)) else [frm]
            return [i - 1 if f == -1 else f for f in frm]

        dest: Dict[int, Tuple[int, int]] = {}
        for j in range(nL):
            if type(layers[j]).__name__ == "Concat":
                off = 0
                for s in src_of(j):
                    if s in dest or type(layers[s]).__name__ == "Concat":
                        raise NotImplementedError("a tensor feeding two Concat layers / nested Concat needs a copy kernel")
                    dest[s] = (j, off)
                    off += Engine._out_channels(layers, s, src_of)
        cat_bufs: Dict[int, ActView] = {}
        outs: List[Optional[ActView]] = [None] * nL

        def out_view(i: int, H: int, W: int, C_: int) -> Optional[ActView]:
            if i not in dest:
                return None
            j, off = dest[i]
            if j not in cat_bufs:
                cat_bufs[j] = self.new_act(H, W, Engine._out_channels(layers, j, src_of))
            return cat_bufs[j].slice(off, C_)

        self.layer_bwd_end: List[int] = []  # len(self.bwd) after top-level layer i was built (its closures end there)
        for i, m in enumerate(layers):
            name = type(m).__name__
            srcs = src_of(i)
            if name in ("Conv", "Focus") and srcs[0] < 0:
                outs[i] = self.stem(m, out_view(i, self.H // 2, self.W // 2, m.conv.out_channels))
                self.layer_bwd_end.append(len(self.bwd))
                continue
            xin = [outs[s] for s in srcs]
            if name == "Conv":
                links = list(m.conv) if isinstance(m.conv, nn.Sequential) else [m.conv]  # Tucker chain: the kxk link strides
                oh, ow = xin[0].H, xin[0].W
                for c in links:
                    k_, s_, p_ = c.kernel_size[0], c.stride[0], c.padding[0]
                    oh, ow = (oh + 2 * p_ - k_) // s_ + 1, (ow + 2 * p_ - k_) // s_ + 1
                outs[i] = self.kindle_conv(m, xin[0], y=out_view(i, oh, ow, links[-1].out_channels))
            elif name == "C3":
                outs[i] = self.c3(m, xin[0], out_view(i, xin[0].H, xin[0].W, m.conv3.conv.out_channels))
            elif name in ("SPP", "SPPF"):
                outs[i] = self.spp(m, xin[0], out_view(i, xin[0].H, xin[0].W, m.conv2.conv.out_channels))
            elif name == "Upsample":
                outs[i] = self.upsample(m, xin[0], out_view(i, 2 * xin[0].H, 2 * xin[0].W, xin[0].c))
            elif name == "Concat":
                outs[i] = cat_bufs[i]
            elif name == "YOLOHead":
                self.head(m, xin)
            else:
                raise NotImplementedError(f"layer {i} ({name}) has no training path")
            self.layer_bwd_end.append(len(self.bwd))

    # ------------------------------------------------------------------------------------------------ gradient buckets
    def plan_grad_buckets(self, n_buckets: int) -> List[Tuple[int, int, int, int]]:
        """Split the backward into `n_buckets` consecutive pieces cut at top-level layer boundaries, of roughly equal
        parameter bytes: [(first closure, last closure + 1, flat gradient offset lo, hi)] in EXECUTION order (last layers
        first). Parameters are laid out in model.parameters() order = layer order, so the gradients a piece completes are
        ONE contiguous range of the flat buffer: it can be all-reduced on a side stream while the next piece still runs
        (the overlap DDP gets from its bucketed hooks, scripts/train/train_model_builder.py:75-78)."""
        layers = list(self.model.model)
        params = list(self.model.parameters())
        off_of = {id(p): o for p, o in zip(params, self.pg_offsets)}
        total = int(self.pg_flat.numel())
        first_off = []  # flat offset where layer i's parameters start (== the next layer's for parameter-free layers)
        nxt = total
        for m in reversed(layers):
            ps = list(m.parameters())
            if ps:
                nxt = min(off_of[id(q)] for q in ps)
            first_off.append(nxt)
        first_off.reverse()
        n_buckets = max(1, min(int(n_buckets), len(layers)))
        cuts, target, hi = [], total / n_buckets, total
        layer_hi = len(layers)
        for i in range(len(layers) - 1, -1, -1):
            lo = first_off[i]
            if (hi - lo >= target and len(cuts) < n_buckets - 1) or i == 0:
                b_lo = self.layer_bwd_end[i - 1] if i > 0 else 0
                cuts.append((b_lo, self.layer_bwd_end[layer_hi - 1], 0 if i == 0 else lo, hi))
                hi, layer_hi = lo, i
        self.bwd_chunks = [c for c in cuts if c[1] > c[0]]
        self.bwd_chunks[-1] = (self.bwd_chunks[-1][0], self.bwd_chunks[-1][1], 0, self.bwd_chunks[-1][3])
        self._gstate = {k: v for k, v in self._gstate.items() if not k.startswith("bwd")}
        return self.bwd_chunks

    # ------------------------------------------------------------------------------------------------ run
    def _run_or_replay(self, key: str, body: Callable[[], None]) -> None:
        st = self._gstate.get(key, 0)
        if not self.use_graph or st == 0:
            body()  # eager (also the warm-up that sets kernel attributes before any capture)
            self._gstate[key] = 1
            return
        if st == 1:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            try:
                # thread-local capture mode: under DDP the NCCL watchdog thread polls its events while this thread captures
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    body()
            except Exception as e:  # capture is an optimisation of the launch path only: keep training, say so loudly
                import warnings

                warnings.warn(f"TrainEngine: CUDA-graph capture of '{key}' failed ({e}); running the kernels eagerly")
                torch.cuda.synchronize()
                self.use_graph = False
                body()
                return
            self._graphs[key] = g
            self._gstate[key] = 2
        self._graphs[key].replay()

    def _forward_body(self) -> None:
        self._sums_flat[:self._sums_used].zero_()
        self._repack_run()
        for r in self.refresh:
            r()
        for f in self.fwd:
            f()
        if self._bn_counters:
            torch._foreach_add_(self._bn_counters, 1)

    def _zero_grad_buffers(self) -> None:
        # Activation-gradient buffers are allocated zeroed and every slice is overwritten by its first writer of the step
        # (_GradSite), so only buffers whose writers overlap partially need a fill -- none in the YOLOv5 graphs; this used
        # to be ~125 fills / 1.8 ms per step at batch 128. The flat parameter gradient accumulates and is always cleared.
        for key in self._needs_zero:
            self._gbuf[key].zero_()
        self.pg_flat.zero_()
        self._sums_flat[:self._sums_used].zero_()

    def _backward_body(self) -> None:
        self._zero_grad_buffers()
        for b in reversed(self.bwd):
            b()

    def forward(self, img: torch.Tensor) -> List[torch.Tensor]:
        # Activations, batch statistics and gradient buffers are ONE static set per engine: a backward is only valid for
        # the most recent forward. The generation counter lets TrainFunction.backward detect fwd/fwd/bwd/bwd patterns
        # (two views summed into one loss, KD double passes) instead of silently differentiating the wrong activations.
        self.generation = getattr(self, "generation", 0) + 1
        if isinstance(img, ResizedInput):
            # multi_scale: the bilinear resize (and prepare_img's / 255) writes the graph's static input directly
            ops.resize_bilinear(img.src, self.static_in, img.pre_scale)
        else:
            self.static_in.copy_(img)
        self._img = self.static_in
        self._repack_sync_table()
        self._run_or_replay("fwd", self._forward_body)
        if getattr(self, "static_grads", False):
            # trainer mode: the loss and its backward run before the next forward (the generation guard enforces it), so
            # the outputs may alias the graph's static buffers -- fresh tensor objects, no copy
            return [o.detach() for o in self.head_out]
        return [o.clone() for o in self.head_out]

    def _backward_chunk(self, k: int) -> None:
        b_lo, b_hi, _, _ = self.bwd_chunks[k]
        if k == 0:
            self._zero_grad_buffers()
        for b in reversed(self.bwd[b_lo:b_hi]):
            b()

    def backward(self, grads: Sequence[torch.Tensor]) -> Dict[int, torch.Tensor]:
        for gl, g_ in zip(self.head_gin, grads):
            g_ = g_.detach()
            if g_.dtype != torch.float32 or not g_.is_contiguous():
                g_ = g_.float().contiguous()
            ops.head_grad_to_nhwc(g_, gl)
        chunks = getattr(self, "bwd_chunks", None)
        if chunks and len(chunks) > 1:
            # bucketed backward: one CUDA graph per piece, `on_grad_bucket(k, lo, hi)` after each (the trainer all-reduces
            # flat gradients [lo, hi) on a side stream while the next piece runs)
            cb = getattr(self, "on_grad_bucket", None)
            for k, (_, _, f_lo, f_hi) in enumerate(chunks):
                self._run_or_replay(f"bwd{k}", lambda k=k: self._backward_chunk(k))
                if cb is not None:
                    cb(k, f_lo, f_hi)
        else:
            self._run_or_replay("bwd", self._backward_body)
        if getattr(self, "static_grads", False):
            flat = self.pg_flat  # the trainer consumes the static buffer in stream order (no copy)
        else:
            flat = self.pg_flat.clone()  # one copy out of the graph's static buffer; the per-parameter gradients are views of it
        self.last_grad_flat = flat
        return {k: flat[v.storage_offset():v.storage_offset() + v.numel()].view(v.shape) for k, v in self.pg.items()}


class TrainFunction(torch.autograd.Function):
    """model(x) in training mode: one autograd node over every parameter of the model."""

    @staticmethod
    def forward(ctx, engine: TrainEngine, x: torch.Tensor, *params: torch.Tensor):
        ctx.engine = engine
        ctx.params = params
        outs = engine.forward(x)
        ctx.generation = engine.generation
        ctx.consumed = False
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts: torch.Tensor):
        eng: TrainEngine = ctx.engine
        if ctx.generation != eng.generation:
            raise RuntimeError(
                "ayolov2_b200 training forward keeps ONE set of activations per (batch, H, W): this backward belongs to "
                f"forward #{ctx.generation} but forward #{eng.generation} has overwritten them. Call loss.backward() before "
                "the next model(x) in train mode (or concatenate the views into one batch).")
        if ctx.consumed:
            raise RuntimeError("ayolov2_b200 training forward cannot be back-propagated twice (retain_graph): the static "
                               "gradient buffers of the first backward have been consumed")
        ctx.consumed = True
        pg = eng.backward([g if g is not None else torch.zeros_like(o) for g, o in zip(gouts, eng.head_out)])
        if getattr(eng, "static_grads", False):
            # trainer mode (ayolov2_b200.trainer.TrainStep): the optimizer consumes the flat gradient buffer directly;
            # no per-parameter .grad tensors are materialised (that would be one copy kernel per parameter)
            return (None, None) + tuple(None for _ in ctx.params)
        grads = tuple(pg.get(id(p)) if p.requires_grad else None for p in ctx.params)
        return (None, None) + grads


class ResizedInput:
    """A training batch that still has to be resized: `src` NCHW uint8 / fp32 on the device, multiplied by `pre_scale` and
    bilinearly resized to `size` (H, W) while it is written into the engine's input (YoloTrainer.multi_scale,
    scripts/train/yolo_trainer.py:223-248, with prepare_img fused in)."""

    def __init__(self, src: torch.Tensor, size: Sequence[int], pre_scale: float = 1.0) -> None:
        assert src.is_cuda and src.dim() == 4 and src.shape[1] == 3 and src.dtype in (torch.uint8, torch.float32)
        self.src, self.size, self.pre_scale = src.contiguous(), (int(size[0]), int(size[1])), float(pre_scale)


def forward_train_resized(model: nn.Module, x: torch.Tensor, size: Sequence[int], pre_scale: float = 1.0) -> List[torch.Tensor]:
    """forward_train(model, F.interpolate(x * pre_scale, size, mode="bilinear", align_corners=False)) without the three
    intermediate fp32 images (scaled copy, interpolated copy, copy into the engine's static input)."""
    inp = ResizedInput(x, size, pre_scale)
    eng = _engine_for(model, x.shape[0], inp.size[0], inp.size[1], x.device, torch.float32, 1.0)
    return list(TrainFunction.apply(eng, inp, *tuple(model.parameters())))


def _engine_for(model: nn.Module, B: int, H: int, W: int, device: torch.device, dtype: torch.dtype, scale: float) -> "TrainEngine":
    cache = model.__dict__.setdefault("_train_engine_cache", {})
    key = (B, H, W, device.index, dtype, float(scale))
    eng = cache.pop(key, None)
    if eng is None:
        while len(cache) >= int(model.__dict__.get("_train_engine_slots", 1)):  # multi_scale training keeps a few shapes
            cache.pop(next(iter(cache)))
        eng = TrainEngine(model, B, H, W, in_dtype=dtype, scale=scale, device=device)
        hook = model.__dict__.get("_train_engine_hook")
        if hook is not None:
            hook(eng)
    cache[key] = eng  # most recently used last
    model.__dict__["_train_engine_last"] = eng
    return eng


def forward_train(model: nn.Module, x: torch.Tensor, scale: float = 1.0) -> List[torch.Tensor]:
    """YOLOModel.forward in training mode (returns the list of (bs, na, ny, nx, no) head outputs) on `x * scale`.
    A uint8 batch with scale = 1/255 is the reference's `prepare_img` (abstract_trainer.py:252-261) fused into the stem's
    space-to-depth pass: no fp32 copy of the images is ever materialised (the trainer's path)."""
    if x.dtype not in (torch.float32, torch.uint8):
        x = x.float()
    B, _, H, W = x.shape
    eng = _engine_for(model, B, H, W, x.device, x.dtype, scale)
    params = tuple(model.parameters())
    return list(TrainFunction.apply(eng, x, *params))
