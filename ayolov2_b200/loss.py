"""Drop-in `ComputeLoss` (scripts/loss/losses.py:168-300) running on the fused CUDA loss kernels.

Same constructor contract (`ComputeLoss(model)` reads model.hyp and the YOLOHead attributes nl/na/nc/anchors/
stride, losses.py:171-221) and same call contract: `loss_fn(preds, targets) -> (loss * bs, cat(lbox, lobj, lcls,
loss).detach())` with `preds` the list of (bs, na, ny, nx, 5+nc) head outputs and `targets` (nt, 6). The returned
loss is differentiable w.r.t. `preds` (torch.autograd.Function whose backward is the analytic gradient kernel).
Supported configuration: plain BCE (fl_gamma == 0, the reference default) or the FocalLoss wrapper (fl_gamma > 0), with or without autobalance.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Tuple

import torch
import torch.nn as nn

from . import _lib


_ACC_BYTES = 5 * 3 * 8  # head of the loss workspace: double acc[AY2_LOSS_MAX_LEVELS][3] = sums of (1 - ciou), cls BCE, obj BCE


def smooth_BCE(eps: float = 0.1) -> Tuple[float, float]:  # noqa: N802 (reference name, losses.py:16)
    return 1.0 - 0.5 * eps, 0.5 * eps


def is_parallel(model: nn.Module) -> bool:
    return type(model) in (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel)


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner: "ComputeLoss", targets: torch.Tensor, *preds: torch.Tensor):
        ctx.owner, ctx.targets = owner, targets
        ctx.balance = list(owner.balance)  # autobalance changes the level weights after the forward; the gradient is of THIS loss
        ctx.save_for_backward(*preds)
        out5 = owner._launch(preds, targets, None, None)
        ctx.mark_non_differentiable(out5)
        return out5[0:1].clone(), out5

    @staticmethod
    def backward(ctx, g_loss: torch.Tensor, _g_items):
        preds = ctx.saved_tensors
        # The kernel writes fp32 at contiguous (cell * no) offsets whatever the dtype / strides of the saved predictions
        # (fp16 / bf16 heads under autocast, permuted views): launch on contiguous fp32 buffers, hand autograd tensors of
        # the predictions' own dtype back.
        grads = [torch.zeros(p.shape, dtype=torch.float32, device=p.device) for p in preds]
        gscale = g_loss.reshape(-1)[:1].float().contiguous()
        ctx.owner._launch(preds, ctx.targets, grads, gscale, balance=ctx.balance)
        return (None, None, *[g if g.dtype == p.dtype else g.to(p.dtype) for g, p in zip(grads, preds)])


class ComputeLoss:
    """Compute YOLO loss on the GPU (fused build_targets + CIoU + BCE, forward and backward)."""

    def __init__(self, model: nn.Module, autobalance: bool = False) -> None:
        hyp: Dict[str, Any] = model.hyp  # type: ignore
        head = model.module.model[-1] if is_parallel(model) else model.model[-1]  # type: ignore
        self.hyp = hyp
        self.cp, self.cn = smooth_BCE(eps=hyp.get("label_smoothing", 0.0))
        self.balance = {3: [4.0, 1.0, 0.4]}.get(head.nl, [4.0, 1.0, 0.25, 0.06, 0.02])
        self.gr, self.autobalance, self.sort_obj_iou = 1.0, bool(autobalance), False
        self.ssi = list(head.stride).index(16) if autobalance else 0  # stride-16 level (losses.py:209)
        self.na, self.nc, self.nl = head.na, head.nc, head.nl
        self.anchors = head.anchors
        self._ws: Dict[Any, torch.Tensor] = {}

    # ---------------------------------------------------------------------------------------------
    def _params(self, preds, nt: int, balance=None) -> _lib.LossParams:
        p = _lib.LossParams()
        p.nl, p.na, p.nc, p.bs, p.nt = self.nl, self.na, self.nc, preds[0].shape[0], nt
        for i, t in enumerate(preds):
            p.ny[i], p.nx[i] = t.shape[2], t.shape[3]
            p.balance[i] = (balance or self.balance)[i]
        h = self.hyp
        p.anchor_t, p.box, p.obj, p.cls = h["anchor_t"], h["box"], h["obj"], h["cls"]
        p.cls_pw, p.obj_pw, p.cp, p.cn = h["cls_pw"], h["obj_pw"], self.cp, self.cn
        p.fl_gamma, p.fl_alpha = float(h.get("fl_gamma", 0.0)), 0.25  # FocalLoss(BCE, gamma) keeps its default alpha (losses.py:71,196)
        return p

    def _launch(self, preds, targets, grads, gscale, balance=None) -> torch.Tensor:
        lib = _lib.load()
        dev = preds[0].device
        if dev.type != "cuda":
            raise RuntimeError("ayolov2_b200.ComputeLoss runs on CUDA tensors only (no CPU fallback)")
        preds = [p if (p.dtype == torch.float32 and p.is_contiguous()) else p.float().contiguous() for p in preds]
        for p_ in preds:
            assert p_.dim() == 5 and p_.shape[1] == self.na and p_.shape[4] == self.nc + 5, p_.shape
        targets = targets.to(dev).float().contiguous()
        anchors = self.anchors.to(dev).float().contiguous()
        nt = targets.shape[0]
        prm = self._params(preds, nt, balance)
        nbytes = lib.ay2_yolo_loss_workspace_bytes(C.byref(prm))
        key = (dev.index, nbytes)
        ws = self._ws.get(key)
        if ws is None:
            self._ws.clear()
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._ws[key] = ws
        out5 = torch.empty(5, dtype=torch.float32, device=dev)
        parr = (C.c_void_p * len(preds))(*[p.data_ptr() for p in preds])
        garr = None
        if grads is not None:
            garr = (C.c_void_p * len(preds))(*[g.data_ptr() for g in grads])
        _lib.check(lib.ay2_yolo_loss(C.byref(prm), parr, garr, targets.data_ptr() if nt else None, anchors.data_ptr(),
                                     _lib.ptr(gscale), ws.data_ptr(), ws.numel(), out5.data_ptr(),
                                     _lib.current_stream_ptr()), "ay2_yolo_loss")
        return out5

    def __call__(self, preds: List[torch.Tensor], targets: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        loss_bs, out5 = _LossFn.apply(self, targets, *preds)
        if self.autobalance:  # losses.py:286-292: per-level objectness loss -> level weights of the NEXT call
            ws = next(iter(self._ws.values()))
            acc = ws[:_ACC_BYTES].view(torch.float64).view(-1, 3)[:self.nl, 2].cpu()  # host sync, like the reference's .item()
            bal = list(self.balance)
            for i, t in enumerate(preds):
                obji = float(acc[i]) / float(t.shape[0] * t.shape[1] * t.shape[2] * t.shape[3])
                bal[i] = bal[i] * 0.9999 + 0.0001 / obji
            self.balance = [x / bal[self.ssi] for x in bal]
        return loss_bs, out5[1:5].detach()
