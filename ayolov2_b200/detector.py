"""Detector: model forward + batched NMS as ONE CUDA graph, plus the end-to-end host pipeline.

This is the fused form of what scripts/utils/train_utils.py:403-472 (YoloValidator.validation_step) does per
batch: prepare_img (uint8 -> float /255, :421,255-260) -> model(imgs) (:436-444) -> non_max_suppression
(:461-469). `detect()` takes HOST images (pinned uint8 or float32 NCHW) and returns the reference's list of
(n_i, 6) tensors; host->device copies, all kernels and the device->host read of the detections are part of
every call. `submit()/collect()` expose the same pipeline asynchronously with two in-flight slots so that the
PCIe copy of batch i+1 overlaps the kernels of batch i.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .engine import Engine
from .nms import nms_device


class Detector:
    def __init__(self, model: nn.Module, batch: int, height: int = 640, width: int = 640, conf_thres: float = 0.25,
                 iou_thres: float = 0.45, multi_label: bool = False, agnostic: bool = False, max_det: int = 300,
                 in_dtype: torch.dtype = torch.uint8, device: Optional[torch.device] = None, want_raw: bool = False,
                 dense_pred: bool = False, fuse_candidates: bool = True, slots: int = 3) -> None:
        scale = 1.0 / 255.0 if in_dtype == torch.uint8 else 1.0
        self.engine = Engine(model, batch, height, width, in_dtype=in_dtype, scale=scale, want_raw=want_raw,
                             device=device, use_graph=False)
        self.device = self.engine.device
        self.conf_thres, self.iou_thres = conf_thres, iou_thres
        self.multi_label, self.agnostic, self.max_det = multi_label, agnostic, max_det
        self.B, self.H, self.W = batch, height, width
        self.in_dtype = in_dtype
        pred = self.engine.pred
        nc = pred.shape[2] - 5
        self._nc = nc
        self.nms_ws = ops.NmsWorkspace(batch, pred.shape[1], pred.shape[2], max_det=max_det,
                                       multi_label=multi_label and nc > 1, device=self.device)
        # fused head: NMS candidates and boxes are decoded straight from the bf16 head logits, the dense
        # (B, 25200, 85) tensor is not written unless `dense_pred` is requested
        self.dense_pred = dense_pred
        eng = self.engine
        self.levels = ops.make_head_levels(eng.head_logits, eng.na, eng.head_strides, eng.head_anchors_px)
        # ... and the candidates themselves are scored by the detect convolutions' epilogues, from the output tile they
        # hold in shared memory (ay2_conv_plan_set_head_candidates): the NMS kernel only sorts and suppresses
        if os.environ.get("AY2_FUSE_CANDIDATES") is not None:  # A/B switch for measurements
            fuse_candidates = os.environ["AY2_FUSE_CANDIDATES"] != "0"
        self.fused_candidates = fuse_candidates and not dense_pred and all(pl.desc.cout_pad <= 256 for pl in eng.head_plans)
        self._arm_candidates()
        # input / output slots of the host pipeline: with `slots` of them, slots - 1 batches can be in flight (the H2D copy
        # of batch i + 1 and the D2H read of batch i - 1 overlap the kernels of batch i)
        self.slots = max(2, int(slots))
        self.dev_in = [torch.zeros((batch, 3, height, width), dtype=in_dtype, device=self.device) for _ in range(self.slots)]
        self.host_out = [torch.zeros((batch, max_det, 6), dtype=torch.float32).pin_memory() for _ in range(self.slots)]
        self.host_cnt = [torch.zeros(batch + 1, dtype=torch.int32).pin_memory() for _ in range(self.slots)]  # [batch] = overflow flag
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ev_h2d = [torch.cuda.Event() for _ in range(self.slots)]
        self.ev_consumed = [torch.cuda.Event() for _ in range(self.slots)]
        self.ev_done = [torch.cuda.Event() for _ in range(self.slots)]
        self._slot = 0
        self._slot_packed: dict = {}  # slot -> (device PackedBatch, fused) for batches submitted through submit_packed
        self.stage: List[Optional[torch.Tensor]] = [None] * self.slots  # raw-image arenas of submit_packed
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._warm = False

    def _arm_candidates(self) -> None:
        if self.fused_candidates:
            eng, p = self.engine, self.nms_ws.p
            p.conf_thres = float(self.conf_thres)
            p.multi_label = int(self.multi_label and self._nc > 1)
            for pl, off in zip(eng.head_plans, eng.head_row_off):
                pl.set_head_candidates(self.nms_ws, eng.na, off)

    def _grow_workspace(self) -> None:
        """Some image had more candidates than the list holds (multi_label at a low conf_thres: up to n * nc): the kernel
        then kept an arbitrary subset. Re-create the workspace with room for every (row, class) pair, re-arm the detect
        convolutions and drop the captured graph (it holds the old workspace's addresses)."""
        n, no = self.engine.pred.shape[1], self.engine.pred.shape[2]
        self.nms_ws = ops.NmsWorkspace(self.B, n, no, max_det=self.max_det, multi_label=self.multi_label and self._nc > 1,
                                       max_candidates=n * max(self._nc, 1), device=self.device)
        self._arm_candidates()
        self._graph, self._warm = None, False

    # ---------------------------------------------------------------------------------------------
    def _body(self) -> None:
        """Everything after the space-to-depth kernel (which reads a per-slot input buffer): convs, head, NMS."""
        eng = self.engine
        if self.fused_candidates:
            self.nms_ws.begin_candidates()
        for s in eng.b.steps:
            if s is eng.b.s2d_step or (not self.dense_pred and s in eng.decode_steps):
                continue
            s()
        if self.fused_candidates:
            self.nms_ws.run_candidates(self.levels, eng.head_logits, self.iou_thres, agnostic=self.agnostic)
        elif self.dense_pred:
            nms_device(eng.pred, self.conf_thres, self.iou_thres, agnostic=self.agnostic, multi_label=self.multi_label,
                       max_det=self.max_det, workspace=self.nms_ws)
        else:
            self.nms_ws.p.multi_label = int(self.multi_label and eng.no - 5 > 1)
            self.nms_ws.run_logits(self.levels, eng.head_logits, self.conf_thres, self.iou_thres, agnostic=self.agnostic)

    def launches_per_step(self) -> int:
        n = len(self.engine.b.steps) - len(self.engine.decode_steps)
        if self.fused_candidates:
            return n + 1  # + the NMS sort/suppress kernel (candidates come from the detect convolutions)
        # + NMS row filter, row scoring and sort/scan kernels (+ the dense decode kernels)
        return n + 3 + (len(self.engine.decode_steps) if self.dense_pred else 0)

    def run_device(self, img: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """img: CUDA NCHW tensor of the detector's dtype. Asynchronous; returns (det [B,max_det,6], count [B])."""
        eng = self.engine
        eng._img = img
        eng.b.s2d_step()
        return self._run_body()

    def _run_body(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """The static part of a step (the stem's space-to-depth input has been written)."""
        if not self._warm:
            self._body()  # first call: plain launches (sets function attributes) ...
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()  # ... then capture the static part once
            with torch.cuda.graph(g):
                self._body()
            self._graph = g
            self._warm = True
        else:
            self._graph.replay()
        return self.nms_ws.out, self.nms_ws.count

    # ---------------------------------------------------------------------------------------------
    def submit(self, host_img: torch.Tensor) -> int:
        """Enqueue one batch from HOST memory (pinned for a truly asynchronous copy). Returns the slot id."""
        k = self._slot
        self._slot = (k + 1) % self.slots
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_consumed[k])  # the s2d kernel that last read this slot has finished
            self.dev_in[k].copy_(host_img, non_blocking=True)
            self.ev_h2d[k].record(self.copy_stream)
        cur.wait_event(self.ev_h2d[k])
        self._slot_packed.pop(k, None)
        det, cnt = self.run_device(self.dev_in[k])
        self.ev_consumed[k].record(cur)
        self.host_out[k].copy_(det, non_blocking=True)
        self.host_cnt[k][:self.B].copy_(cnt, non_blocking=True)
        self.host_cnt[k][self.B:].copy_(self.nms_ws.overflow, non_blocking=True)
        self.ev_done[k].record(cur)
        return k

    def submit_packed(self, pb, fused: bool = True) -> int:
        """Enqueue one batch of LOADED images (`data_loader.pack_batch`: ragged HWC BGR uint8 + geometry table in one host
        arena). The letterbox / channel flip / collate of scripts/data_loader/data_loader.py:380-393,461-477 runs on the
        device: one H2D copy of the raw bytes, then one kernel that writes the stem's space-to-depth input directly
        (`fused`; prepare_img's /255 included) or the uint8 NCHW slot followed by the usual first step. Returns the slot id;
        `pb.shapes` holds what `scale_coords` needs."""
        import dataclasses

        assert pb.batch == self.B and tuple(pb.out_shape) == (self.H, self.W) and self.in_dtype == torch.uint8
        k = self._slot
        self._slot = (k + 1) % self.slots
        n, need = pb.arena.numel(), pb.arena.numel() + pb.scratch_bytes  # + device scratch of the decode-side resize
        if self.stage[k] is None or self.stage[k].numel() < need:
            torch.cuda.synchronize(self.device)  # a (rare) growth must not free bytes a running kernel still reads
            self.stage[k] = torch.empty(max(need, self.B * (3 * self.H * self.W + 64)), dtype=torch.uint8, device=self.device)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_consumed[k])
            dst = self.stage[k][:need]
            dst[:n].copy_(pb.arena, non_blocking=True)
            self.ev_h2d[k].record(self.copy_stream)
        cur.wait_event(self.ev_h2d[k])
        dpb = dataclasses.replace(pb, arena=dst)
        eng = self.engine
        fused = fused and not eng.b.x3
        self._slot_packed[k] = (dpb, fused)  # the slot's input, should collect() have to redo the batch (candidate overflow)
        det, cnt = self._run_packed(dpb, fused, k)
        self.ev_consumed[k].record(cur)
        self.host_out[k].copy_(det, non_blocking=True)
        self.host_cnt[k][:self.B].copy_(cnt, non_blocking=True)
        self.host_cnt[k][self.B:].copy_(self.nms_ws.overflow, non_blocking=True)
        self.ev_done[k].record(cur)
        return k

    def _run_packed(self, dpb, fused: bool, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        eng = self.engine
        if fused:
            dpb.to_space_to_depth(eng.b.s2d_view, eng.b.s2d_scale, x_offset=1)
            return self._run_body()
        dpb.to_device(out=self.dev_in[k])
        return self.run_device(self.dev_in[k])

    def collect(self, k: int) -> List[torch.Tensor]:
        """Wait for slot k and return the reference-style list of (n_i, 6) tensors (host memory)."""
        self.ev_done[k].synchronize()
        if int(self.host_cnt[k][self.B]):
            # candidate-list overflow: redo this batch (still in its device slot) with a workspace that cannot overflow
            torch.cuda.synchronize(self.device)
            self._grow_workspace()
            packed = self._slot_packed.get(k)
            det, cnt = self._run_packed(packed[0], packed[1], k) if packed else self.run_device(self.dev_in[k])
            self.host_out[k].copy_(det)
            self.host_cnt[k][:self.B].copy_(cnt)
            self.host_cnt[k][self.B] = 0
        counts = self.host_cnt[k][:self.B].tolist()
        out = self.host_out[k]
        return [out[i, :c].clone() for i, c in enumerate(counts)]

    def detect(self, host_img: torch.Tensor) -> List[torch.Tensor]:
        """Synchronous end-to-end call: H2D, forward, NMS, D2H."""
        return self.collect(self.submit(host_img))

    @property
    def flops_per_step(self) -> float:
        return self.engine.b.flops

    @property
    def act_bytes_per_step(self) -> float:
        return self.engine.b.act_bytes
