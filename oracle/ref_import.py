"""Import the UNMODIFIED reference loss / NMS code from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY — used to pin oracle/*.py against the reference's own functions and to generate
the golden fixtures in tests/golden/ (tests/golden/make_golden.py). /root/reference does not exist on the
GPU box; callers must check `available()` first.

The reference modules import three packages that are absent here: `matplotlib`,
`matplotlib.pyplot` (plot helpers only) and `kindle` (type annotations only on this path). They are stubbed
with empty modules. `ComputeLoss.build_targets` (scripts/loss/losses.py:385) calls `Tensor.clamp_` with a
float *tensor* bound on a long tensor, which torch >= 1.10 rejects (the reference pins torch 1.9.1,
environment.yml:27); `clamp_compat()` converts tensor bounds to python numbers for the duration of a call —
the semantics (clamp to [0, n-1]) are unchanged.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REF_ROOT = os.environ.get("AY2_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "scripts", "loss", "losses.py"))


def _stub(name: str) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__dict__["__ay2_stub__"] = True
        sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns a namespace with ComputeLoss, non_max_suppression, bbox_iou, box_iou, batched_nms, xywh2xyxy."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True  # /root/reference is read-only
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    if not hasattr(mpl, "use"):
        mpl.use = lambda *a, **k: None
    kindle = sys.modules.get("kindle")
    if kindle is None:
        kindle = _stub("kindle")
    for attr in ("YOLOModel", "Model"):
        if not hasattr(kindle, attr):
            setattr(kindle, attr, type(attr, (), {}))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from scripts.loss.losses import ComputeLoss  # type: ignore
    from scripts.utils.general import xywh2xyxy  # type: ignore
    from scripts.utils.metrics import bbox_iou, box_iou, non_max_suppression  # type: ignore
    from scripts.utils.nms import batched_nms  # type: ignore

    _loaded.update(ComputeLoss=ComputeLoss, non_max_suppression=non_max_suppression, bbox_iou=bbox_iou,
                   box_iou=box_iou, batched_nms=batched_nms, xywh2xyxy=xywh2xyxy)
    return types.SimpleNamespace(**_loaded)


@contextlib.contextmanager
def clamp_compat():
    """Make `long_tensor.clamp_(0, float_tensor)` work on torch >= 1.10 (losses.py:385)."""
    import torch

    orig = torch.Tensor.clamp_

    def clamp_(self, min=None, max=None):  # noqa: A002
        if isinstance(min, torch.Tensor):
            min = min.item()
        if isinstance(max, torch.Tensor):
            max = max.item()
        if not self.is_floating_point():
            min = None if min is None else int(min)
            max = None if max is None else int(max)
        return orig(self, min, max)

    torch.Tensor.clamp_ = clamp_
    try:
        yield
    finally:
        torch.Tensor.clamp_ = orig


def load_data_loader():
    """The reference's scripts/data_loader/data_loader.py module, unmodified (for LoadImages._letterbox / collate_fn,
    LoadImagesAndLabels.collate_fn). Needs cv2 (installed in the build container) and one more stub: `p_tqdm`
    (progress-bar helper of the label cache, not on this path)."""
    load()
    pt = _stub("p_tqdm")
    if not hasattr(pt, "p_map"):
        pt.p_map = None
    import scripts.data_loader.data_loader as dl  # type: ignore

    return dl


class _AnyAttr(types.ModuleType):
    """A stub module whose every attribute is an empty class (for `from albumentations import DualTransform` etc.)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def load_kd_trainer():
    """The reference's scripts/train/kd_trainer.py, unmodified (SoftTeacherTrainer.prepare_labels_for_augmention / filter_invalid
    are plain functions of their arguments). Two more absent packages are stubbed: `albumentations` (strong augmentation, not
    on this path) and `wandb`."""
    load_data_loader()
    for name in ("albumentations", "wandb"):
        if name not in sys.modules:
            sys.modules[name] = _AnyAttr(name)
    import scripts.train.kd_trainer as kd  # type: ignore

    return kd
