"""Drop-in training step: what `YoloTrainer.training_step` (scripts/train/yolo_trainer.py:289-358) and the pieces of
`YoloTrainer.__init__` / `_init_optimizer` / `warmup` / `multi_scale` it relies on (:84-127, 140-248) do per batch, on the
libay2 kernels.

    step = TrainStep(model, hyp, batch_size=128, batches_per_epoch=nb, epochs=300)
    loss = step.training_step((imgs, labels, paths, shapes), batch_idx, epoch)

Same schedule semantics as the reference:
  * nominal batch 64: `accumulate = max(round(64 / batch_size), 1)`, weight decay scaled by batch_size * accumulate / 64;
    during warm-up `accumulate` ramps from 1 (yolo_trainer.py:200-204);
  * three parameter groups -- BatchNorm weights (no decay), other weights (decay), biases (no decay) -- with the warm-up
    ramps of :206-221 (bias lr falls from warmup_bias_lr, the others rise from 0, momentum rises from warmup_momentum) and
    the cosine / linear epoch schedule of :124-138 (`lr_function`);
  * `loss *= world_size` under DDP (:325-326) because the gradient exchange averages;
  * optimizer step + EMA (torch_utils.py:377-416, decay ramp 0.9999 * (1 - exp(-updates / 2000))) every `accumulate` batches;
  * `multi_scale` (:223-248): a random size in [0.5, 1.5] x img_size rounded to the grid, bilinear resize on the device.
What differs is the machinery: the forward / backward are the CUDA graphs of `train_engine.TrainEngine`; parameters,
momenta and the EMA live in ONE flat fp32 buffer each, so the optimizer + EMA is one fused launch
(`ay2_sgd_ema_step_groups`) whatever the group interleaving; activations are bf16 with fp32 accumulation, so the
reference's fp16 `GradScaler` has nothing to scale (`scaler` is None); and under DDP the flat gradient is all-reduced in
a few contiguous buckets on a side stream WHILE the rest of the backward runs (NCCL over NVLink), instead of through
autograd hooks.
"""
from __future__ import annotations

import math
import random
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dist_utils as du
from . import ops
from .loss import ComputeLoss
from .train_engine import forward_train, forward_train_resized

NOMINAL_BATCH = 64  # yolo_trainer.py:88


def parameter_groups(model: nn.Module) -> Dict[int, int]:
    """id(parameter) -> group: 0 = BatchNorm weights, 1 = other weights (decayed), 2 = biases (yolo_trainer.py:149-160)."""
    group: Dict[int, int] = {}
    for mod in model.modules():
        b = getattr(mod, "bias", None)
        if isinstance(b, torch.Tensor):
            group[id(b)] = 2
        w = getattr(mod, "weight", None)
        if isinstance(w, torch.Tensor):
            group[id(w)] = 0 if isinstance(mod, nn.BatchNorm2d) else 1
    return group


def lr_function(epoch: int, epochs: int, lrf: float, linear: bool = False) -> float:
    """Multiplier of the initial learning rate at `epoch` (yolo_trainer.py:124-138)."""
    if linear:
        return (1 - epoch / (epochs - 1)) * (1.0 - lrf) + lrf
    return ((1 + math.cos(epoch * math.pi / epochs)) / 2) * (1 - lrf) + lrf


def warmup_state(ni: int, num_warmups: float, epoch_factor: float, lr0: float, hyp: Dict[str, Any], batch_size: int
                 ) -> Tuple[int, List[float], float]:
    """(accumulate, [lr of group 0, 1, 2], momentum) at integrated batch `ni` <= num_warmups (yolo_trainer.py:194-221)."""
    xs = [0, num_warmups]
    accumulate = int(max(1, np.interp(ni, xs, [1, NOMINAL_BATCH / batch_size]).round()))
    lrs = [float(np.interp(ni, xs, [hyp["warmup_bias_lr"] if j == 2 else 0.0, lr0 * epoch_factor])) for j in range(3)]
    momentum = float(np.interp(ni, xs, [hyp["warmup_momentum"], hyp["momentum"]]))
    return accumulate, lrs, momentum


class TrainStep:
    def __init__(self, model: nn.Module, hyp: Dict[str, Any], batch_size: int, batches_per_epoch: int, epochs: int = 300,
                 img_size: int = 640, multi_scale: bool = False, linear_lr: bool = False, use_ema: bool = True,
                 label_smoothing: float = 0.0, grad_buckets: int = 4, device: Optional[torch.device] = None) -> None:
        """batch_size: the GLOBAL batch (cfg_train["batch_size"]); each rank feeds batch_size // world_size images."""
        if not torch.cuda.is_available():
            raise RuntimeError("ayolov2_b200.TrainStep needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.rank, self.local_rank, self.world = du.env_ranks()
        self.device = torch.device(device) if device is not None else torch.device("cuda", self.local_rank)
        self.model = model.to(self.device).train()
        self.hyp = dict(hyp)
        self.hyp["label_smoothing"] = label_smoothing       # yolo_trainer.py:83
        opt = dict(self.hyp.get("optimizer_params", {}))
        self.lr0 = float(opt.get("lr", 0.01))
        self.hyp.setdefault("momentum", float(opt.get("momentum", 0.937)))
        self.nesterov = bool(opt.get("nesterov", True))
        self.batch_size, self.epochs, self.nb = int(batch_size), int(epochs), int(batches_per_epoch)
        self.linear_lr, self.use_multi_scale = linear_lr, multi_scale
        self.accumulate = max(round(NOMINAL_BATCH / self.batch_size), 1)                               # :89
        self.weight_decay = float(self.hyp.get("weight_decay", 5e-4)) * self.batch_size * self.accumulate / NOMINAL_BATCH  # :143-146
        self.num_warmups = max(round(float(self.hyp.get("warmup_epochs", 3.0)) * self.nb), 1e3)        # :105-107
        stride = int(max(float(s) for s in self.model.stride))
        self.grid_size = stride
        self.img_size = int(math.ceil(img_size / stride) * stride)                                     # check_img_size
        self.model.hyp = self.hyp
        self.loss = ComputeLoss(self.model)
        self.scaler = None  # bf16 storage + fp32 accumulation: no loss scaling (the reference's GradScaler guards fp16)
        self.momentum = float(self.hyp["momentum"])
        self.lrs = [self.lr0] * 3
        self.ema_updates = 0
        self.use_ema = use_ema and self.rank == 0  # the reference keeps the EMA on RANK in (-1, 0) only
        self.grad_buckets = int(grad_buckets)
        self._flat_ready = False
        self._side = torch.cuda.Stream(device=self.device) if self.world > 1 else None
        self._micro = 0
        self.mloss: Optional[torch.Tensor] = None
        self.skip_exchange = False  # measurement aid (bench.py): run the step without the gradient exchange

    # ------------------------------------------------------------------------------------------------ flat state
    def _setup_flat(self, eng) -> None:
        """Parameters become views of one flat fp32 buffer in the engine's gradient layout; momenta, EMA, groups follow."""
        params = list(self.model.parameters())
        groups = parameter_groups(self.model)
        self.flat_p = torch.zeros_like(eng.pg_flat)
        self.flat_group = torch.zeros(eng.pg_flat.numel(), dtype=torch.uint8, device=self.device)
        for p, o in zip(params, eng.pg_offsets):
            self.flat_p[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + p.numel()].view(p.shape)
            self.flat_group[o:o + p.numel()] = groups.get(id(p), 1)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_ema = self.flat_p.clone() if self.use_ema else None
        self.ema_buffers = ({k: v.detach().clone() for k, v in self.model.named_buffers() if v.dtype.is_floating_point}
                            if self.use_ema else {})
        self.flat_acc = torch.zeros_like(self.flat_p) if self.accumulate_max() > 1 else None
        self._bucket_events: List[torch.cuda.Event] = []
        self._flat_ready = True
        self._configure_engine(eng)
        self.model.__dict__["_train_engine_hook"] = self._configure_engine  # engines of later input shapes (multi_scale)
        if self.use_multi_scale:
            self.model.__dict__["_train_engine_slots"] = 4

    def _configure_engine(self, eng) -> None:
        eng.static_grads = True
        if self.world > 1 and self.grad_buckets > 1:
            eng.plan_grad_buckets(self.grad_buckets)
            eng.on_grad_bucket = self._on_grad_bucket

    def accumulate_max(self) -> int:
        return max(self.accumulate, int(max(1, round(NOMINAL_BATCH / self.batch_size))))

    def ema_state_dict(self) -> Dict[str, torch.Tensor]:
        """The EMA model's floating-point state (parameters from the flat buffer, buffers averaged separately)."""
        assert self.use_ema and self._flat_ready
        out = {}
        eng = self._engine()
        for (name, p), o in zip(self.model.named_parameters(), eng.pg_offsets):
            out[name] = self.flat_ema[o:o + p.numel()].view(p.shape)
        out.update(self.ema_buffers)
        return out

    def _engine(self):
        return self.model.__dict__["_train_engine_last"]

    # ------------------------------------------------------------------------------------------------ schedule
    def warmup(self, ni: int, epoch: int) -> None:
        f = lr_function(epoch, self.epochs, float(self.hyp.get("lrf", 0.1)), self.linear_lr)
        self.accumulate, self.lrs, self.momentum = warmup_state(ni, self.num_warmups, f, self.lr0, self.hyp, self.batch_size)

    def set_epoch_lr(self, epoch: int) -> None:
        """scheduler.step() of the reference's LambdaLR at the start of `epoch` (all three groups share lr0)."""
        f = lr_function(epoch, self.epochs, float(self.hyp.get("lrf", 0.1)), self.linear_lr)
        self.lrs = [self.lr0 * f] * 3
        self.momentum = float(self.hyp["momentum"])

    def multi_scale_shape(self, hw: Sequence[int]) -> Optional[List[int]]:
        """The random target size of `multi_scale` (:233-245; one draw from `random` per step, like the reference), or None
        when the batch keeps its size."""
        g = self.grid_size
        sz = random.randrange(int(self.img_size * 0.5), int(self.img_size * 1.5 + g)) // g * g
        sf = sz / max(hw)
        return [math.ceil(x * sf / g) * g for x in hw] if sf != 1 else None

    def multi_scale(self, imgs: torch.Tensor) -> torch.Tensor:
        """Reference form on a prepared float batch (kept for callers that hold one; the step itself resizes while it fills
        the engine's input, `forward_train_resized`)."""
        new_shape = self.multi_scale_shape(imgs.shape[2:])
        if new_shape is not None:
            imgs = F.interpolate(imgs, size=new_shape, mode="bilinear", align_corners=False)
        return imgs

    @staticmethod
    def prepare_img(imgs: torch.Tensor, device: torch.device) -> torch.Tensor:
        """abstract_trainer.py:252-261: uint8 0-255 -> float 0-1 on the device."""
        return imgs.to(device, non_blocking=True).float() / 255.0

    # ------------------------------------------------------------------------------------------------ DDP exchange
    def _on_grad_bucket(self, k: int, lo: int, hi: int) -> None:
        """Called by the engine after backward piece k was enqueued: its gradients [lo, hi) of the flat buffer are final."""
        if self._exchange_now:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._side):
                self._side.wait_event(ev)
                torch.distributed.all_reduce(self._engine().pg_flat[lo:hi])
            done = torch.cuda.Event()
            done.record(self._side)
            self._bucket_events.append(done)

    # ------------------------------------------------------------------------------------------------ the step
    def training_step(self, train_batch: Sequence[Any], batch_idx: int, epoch: int) -> torch.Tensor:
        ni = batch_idx + self.nb * epoch
        if ni <= self.num_warmups:
            self.warmup(ni, epoch)
        imgs, labels = train_batch[0], train_batch[1]
        u8 = imgs.dtype == torch.uint8 and self.model.training
        new_shape = self.multi_scale_shape(imgs.shape[2:]) if self.use_multi_scale else None
        fused_u8 = u8 and new_shape is None
        if u8:
            imgs = imgs.to(self.device, non_blocking=True)  # / 255 happens inside the stem's space-to-depth / the resize pass
        else:
            imgs = self.prepare_img(imgs, self.device) if imgs.dtype == torch.uint8 else imgs.to(self.device).float()
        labels = labels.to(self.device)
        boundary = ni % self.accumulate == 0
        self._exchange_now = self.world > 1 and boundary and self.accumulate == 1 and self._flat_ready and not self.skip_exchange
        self._bucket_events = []
        if fused_u8:
            pred = forward_train(self.model, imgs, scale=1.0 / 255.0)
        elif new_shape is not None and self.model.training:
            # multi_scale (:223-248): prepare_img's / 255, the bilinear resize and the copy into the engine's input are one pass
            pred = forward_train_resized(self.model, imgs, new_shape, pre_scale=1.0 / 255.0 if u8 else 1.0)
        else:
            if new_shape is not None:
                imgs = F.interpolate(imgs, size=new_shape, mode="bilinear", align_corners=False)
            pred = self.model(imgs)
        eng = self._engine()
        if not self._flat_ready:
            self._setup_flat(eng)  # first call: the engine (and the flat layout) exists now; this step exchanges unbucketed
        loss, loss_items = self.loss(pred, labels)
        if self.world > 1:
            loss = loss * self.world                                                                   # :325-326
        loss.backward()
        g = eng.last_grad_flat
        if self.flat_acc is not None and self.accumulate > 1:
            self.flat_acc.add_(g)
            g = self.flat_acc
        if boundary:
            cur = torch.cuda.current_stream(self.device)
            if self.world > 1 and not self.skip_exchange:
                if self._bucket_events:
                    for ev in self._bucket_events:  # the bucketed exchange ran beside the backward: wait for its tail only
                        cur.wait_event(ev)
                else:
                    torch.distributed.all_reduce(g)
            self._optimizer_step(g)
            if self.flat_acc is not None:
                self.flat_acc.zero_()
        for p in self.model.parameters():
            p.grad = None
        li = loss_items.detach()
        self.mloss = li if self.mloss is None else (self.mloss * batch_idx + li) / (batch_idx + 1)      # :270
        return loss[0] if loss.dim() else loss

    def _optimizer_step(self, g: torch.Tensor) -> None:
        d = 0.0
        if self.use_ema:
            self.ema_updates += 1
            d = 0.9999 * (1 - math.exp(-self.ema_updates / 2000))                                      # torch_utils.py:399-401
        ops.sgd_ema_step_groups(self.flat_p, g, self.flat_m, self.flat_ema, self.flat_group, self.lrs,
                                [0.0, self.weight_decay, 0.0], self.momentum, self.nesterov, d,
                                grad_scale=1.0 / self.world)  # DDP averages: sum all-reduce, then 1 / world inside the kernel
        if self.use_ema and self.ema_buffers:  # BatchNorm running statistics are part of the EMA state as well (:411-416)
            cur = dict(self.model.named_buffers())
            keys = list(self.ema_buffers)
            torch._foreach_mul_([self.ema_buffers[k] for k in keys], d)
            torch._foreach_add_([self.ema_buffers[k] for k in keys], [cur[k].detach() for k in keys], alpha=1.0 - d)
