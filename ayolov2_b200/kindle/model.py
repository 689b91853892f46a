"""kindle.model — YOLOModel / ModelParser, API-compatible with the (un-vendored) `kindle` package the reference
builds its models with (train.py:137, val.py:16, scripts/utils/torch_utils.py:259, decompose_model.py:17).

yaml grammar (res/configs/model/*.yaml; rules verified by reproducing every published parameter count,
SURVEY.md §8a):  rows `[from, repeat, Module, args(, kwargs)]`; out-channels `ceil(c * width_multiple / 8) * 8`;
`repeat > 1` -> `max(round(repeat * depth_multiple), 1)` inner bottlenecks; YOLOHead channels are not scaled.

Execution is NOT done by these modules' PyTorch ops: `YOLOModel.forward` hands CUDA inputs to
ayolov2_b200.engine (hand-written sm_100a kernels behind the C-ABI). There is no CPU fallback.
"""
from __future__ import annotations

import math
import time
from copy import deepcopy
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import yaml

from . import modules as M


def make_divisible(v: float, divisor: int = 8) -> int:
    return int(math.ceil(v / divisor) * divisor)


class ModelParser:
    """Builds the nn.Sequential from a model config (attribute names follow the pickled fixture)."""

    def __init__(self, cfg: Union[str, Dict[str, Any]] = "./model_configs/show_case.yaml", verbose: bool = False) -> None:
        self.verbose = verbose
        if isinstance(cfg, dict):
            self.cfg = cfg
        else:
            with open(cfg) as f:
                self.cfg = yaml.load(f, yaml.FullLoader)
        self.input_size = self.cfg.get("input_size", None)
        self.custom_module_paths = self.cfg.get("custom_module_paths", None)
        self.in_channel = self.cfg["input_channel"]
        self.depth_multiply = self.cfg["depth_multiple"]
        self.width_multiply = self.cfg["width_multiple"]
        self.channel_divisor = self.cfg.get("channel_divisor", 8)
        self.backbone_cfg = self.cfg["backbone"]
        self.head_cfg = self.cfg.get("head", [])
        self.model, self.output_save, self.scales = self._parse_model()

    def log(self, msg: str) -> None:
        if self.verbose:
            print(msg)

    def _ch(self, c: int) -> int:
        return make_divisible(c * self.width_multiply, self.channel_divisor)

    def _parse_model(self) -> Tuple[nn.Sequential, List[int], List[float]]:
        layers: List[nn.Module] = []
        out_ch: List[int] = []       # channels of every layer output
        scales: List[float] = []     # spatial down-scale of every layer output w.r.t. the input
        save: List[int] = []
        rows = list(self.backbone_cfg) + list(self.head_cfg)
        self.log(f"{'idx':>3} | {'from':>12} | {'n':>2} | {'params':>9} | {'module':>14} | arguments")
        for i, row in enumerate(rows):
            frm, repeat, name, args = row[0], row[1], row[2], list(row[3])
            kwargs = dict(row[4]) if len(row) > 4 else {}
            n_rep = max(round(repeat * self.depth_multiply), 1) if repeat > 1 else repeat
            src = frm if isinstance(frm, list) else [frm]
            src_abs = [(i - 1 if f == -1 else (f if f >= 0 else i + f)) for f in src]
            cin = [self.in_channel if s < 0 else out_ch[s] for s in src_abs]
            sin = [1.0 if s < 0 else scales[s] for s in src_abs]
            act = kwargs.get("activation", "ReLU")
            if name == "Conv":
                c2 = self._ch(args[0])
                k = args[1] if len(args) > 1 else 1
                s = args[2] if len(args) > 2 else 1
                p = args[3] if len(args) > 3 else None
                m: nn.Module = M.Conv(cin[0], c2, k, s, p, activation=act)
                co, sc = c2, sin[0] * s
            elif name == "Focus":
                c2 = self._ch(args[0])
                k = args[1] if len(args) > 1 else 1
                m = M.Focus(cin[0], c2, k, activation=act)
                co, sc = c2, sin[0] * 2
            elif name in ("C3", "BottleneckCSP"):
                c2 = self._ch(args[0])
                shortcut = args[1] if len(args) > 1 else True
                cls = M.C3 if name == "C3" else M.BottleneckCSP
                m = cls(cin[0], c2, n_repeat=n_rep, shortcut=shortcut, activation=act)
                co, sc = c2, sin[0]
            elif name == "SPP":
                c2 = self._ch(args[0])
                m = M.SPP(cin[0], c2, tuple(args[1]) if len(args) > 1 else (5, 9, 13), activation=act)
                co, sc = c2, sin[0]
            elif name == "SPPF":
                c2 = self._ch(args[0])
                m = M.SPPF(cin[0], c2, args[1] if len(args) > 1 else 5, activation=act)
                co, sc = c2, sin[0]
            elif name == "UpSample":
                size = args[0] if len(args) > 0 else None
                factor = args[1] if len(args) > 1 else 2
                mode = args[2] if len(args) > 2 else "nearest"
                m = nn.Upsample(size=size, scale_factor=float(factor) if factor is not None else None, mode=mode)
                co, sc = cin[0], sin[0] / float(factor)
            elif name == "Concat":
                m = M.Concat(args[0] if args else 1)
                co, sc = sum(cin), sin[0]
            elif name == "YOLOHead":
                m = M.YOLOHead(cin, args[0], args[1], *(args[2:3]))
                m.stride = torch.tensor(sin)
                co, sc = m.no * m.na, sin[0]
            else:
                raise NotImplementedError(
                    f"module {name!r} (row {i}) is outside the YOLO detection hot path this package implements")
            m.n_params = sum(p.numel() for p in m.parameters())  # type: ignore
            m.name, m.module_idx, m.from_idx = name, i, frm  # type: ignore
            layers.append(m)
            out_ch.append(co)
            scales.append(sc)
            save.extend(f for f in (frm if isinstance(frm, list) else [frm]) if f != -1)
            self.log(f"{i:3d} | {str(frm):>12} | {n_rep:2d} | {m.n_params:9,d} | {name:>14} | {args}")
        return nn.Sequential(*layers), list(dict.fromkeys(save)), scales


class YOLOModel(nn.Module):
    """`YOLOModel(cfg_dict_or_yaml_path, verbose=False)` — see module docstring for the contract."""

    def __init__(self, cfg: Union[str, Dict[str, Any]] = "./model_configs/show_case.yaml", verbose: bool = False,
                 init_bias: bool = False) -> None:
        super().__init__()
        self.model_parser = ModelParser(cfg=cfg, verbose=verbose)
        self.model = self.model_parser.model
        self.output_save = self.model_parser.output_save
        head = self.model[-1]
        if isinstance(head, M.YOLOHead):
            self.stride = head.stride.clone()
            head.anchors /= head.stride.to(head.anchors.device).view(-1, 1, 1)
            self.initialize_biases = head.initialize_biases
            if init_bias:
                self.initialize_biases()
        else:
            self.stride = torch.tensor([32.0])
        for m in self.modules():  # ultralytics/kindle convention for the detection models
            if isinstance(m, nn.BatchNorm2d):
                m.eps, m.momentum = 1e-3, 0.03
        self._engine_cache: Dict[Any, Any] = {}

    # -------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, augment: bool = False, profile: bool = False):
        if not x.is_cuda:
            raise RuntimeError("ayolov2_b200 YOLOModel runs on CUDA (sm_100a) tensors only; there is no CPU fallback "
                               "(the CPU restatement lives in oracle/ and is test infrastructure)")
        if self.training:
            from ..train_engine import forward_train

            return forward_train(self, x)
        from ..engine import forward_model

        return forward_model(self, x)

    def invalidate_engine(self) -> None:
        self._engine_cache.clear()
        self.__dict__.pop("_train_engine_cache", None)

    def _apply(self, fn, *a, **k):  # .to()/.half()/.float()/.cuda() move or recast parameters
        self.__dict__.get("_engine_cache", {}).clear()
        self.__dict__.pop("_train_engine_cache", None)
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.__dict__.get("_engine_cache", {}).clear()
        return super().load_state_dict(*a, **k)

    def __deepcopy__(self, memo):
        cache = self.__dict__.pop("_engine_cache", {})
        tcache = self.__dict__.pop("_train_engine_cache", None)
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                new.__dict__[k] = deepcopy(v, memo)
            new.__dict__["_engine_cache"] = {}
        finally:
            self.__dict__["_engine_cache"] = cache
            if tcache is not None:
                self.__dict__["_train_engine_cache"] = tcache
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engine_cache"] = {}
        d.pop("_train_engine_cache", None)
        return d

    def __setstate__(self, d):
        super().__setstate__(d)
        self.__dict__.setdefault("_engine_cache", {})

    # -------------------------------------------------------------------------------------------------
    def fuse(self) -> "YOLOModel":
        """Fold every BatchNorm into its convolution (val.py:331). Parameter count drops by sum(C_bn)... x2
        (gamma, beta) minus the new conv biases (tests/test_tensor_decomposition.py:47: 7,276,605 -> 7,266,973)."""
        for m in self.modules():
            if isinstance(m, (M.Conv, M.Focus)) and isinstance(m.batch_norm, nn.modules.batchnorm._BatchNorm):
                fuse_conv_and_bn(m)
        self.invalidate_engine()
        return self

    def export(self, verbose: bool = False) -> "YOLOModel":
        """Fuse and switch the head to export behaviour: forward returns a tuple whose [0] is the concatenated
        prediction tensor (tests/test_model_convert.py:33-44)."""
        self.fuse()
        self.eval()
        self.__dict__["_export"] = True
        return self

    def profile(self, input_size: Tuple[int, int] = (128, 128), batch_size: int = 1, n_run: int = 100,
                verbose: bool = True, **_: Any):
        """Whole-model timing with CUDA events (kindle profiles per layer; here one fused graph is timed)."""
        dev = next(self.parameters()).device
        x = torch.rand((batch_size, self.model_parser.in_channel, *input_size), device=dev)
        was_training = self.training
        self.eval()
        for _i in range(3):
            self(x)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _i in range(n_run):
            self(x)
        end.record()
        torch.cuda.synchronize()
        self.train(was_training)
        total = start.elapsed_time(end) / 1000.0

        class _Profile:
            total_run_time = total
            running_time = total / n_run
            n_run_ = n_run

        if verbose:
            print(f"profile: {n_run} runs of {tuple(x.shape)} in {total:.4f}s ({total / n_run * 1e3:.3f} ms/run)")
        return _Profile()


def fuse_conv_and_bn(m: nn.Module) -> None:
    """In-place fold of m.batch_norm into m.conv (conv may be a Tucker nn.Sequential: the BN goes into its last
    1x1, which already carries the bias slot, scripts/tensor_decomposition/decomposition.py:395-412)."""
    bn = m.batch_norm
    conv = m.conv[-1] if isinstance(m.conv, nn.Sequential) else m.conv
    with torch.no_grad():
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w = conv.weight * scale.view(-1, 1, 1, 1)
        b0 = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
        b = bn.bias + (b0 - bn.running_mean) * scale
        conv.weight = nn.Parameter(w.to(conv.weight.dtype), requires_grad=conv.weight.requires_grad)
        conv.bias = nn.Parameter(b.to(conv.weight.dtype), requires_grad=conv.weight.requires_grad)
    m.batch_norm = nn.Identity()


class Model(nn.Module):
    """`kindle.Model` (generic classifier builder) is used only by the representation-learning scripts
    (train_repr.py) which are outside the detection hot path."""

    def __init__(self, *a: Any, **k: Any) -> None:
        super().__init__()
        raise NotImplementedError("kindle.Model (representation-learning heads) is out of scope for ayolov2_b200")
