"""`import kindle` shim: the reference's entry points (train.py:12, val.py:16, decompose_model.py:17, ...) and its
whole-module checkpoints (pickles referencing kindle.model.YOLOModel, kindle.modules.conv.Conv, ...) resolve to
the B200-native implementation in ayolov2_b200.kindle when the repo root is on sys.path."""
import sys as _sys

from ayolov2_b200.kindle import Model, ModelParser, YOLOModel, model, modules  # noqa: F401
from ayolov2_b200.kindle.modules import activation, bottleneck, concat, conv, poolings, yolo_head  # noqa: F401

for _name, _mod in {
    "kindle.model": model,
    "kindle.modules": modules,
    "kindle.modules.activation": activation,
    "kindle.modules.bottleneck": bottleneck,
    "kindle.modules.concat": concat,
    "kindle.modules.conv": conv,
    "kindle.modules.poolings": poolings,
    "kindle.modules.yolo_head": yolo_head,
}.items():
    _sys.modules.setdefault(_name, _mod)

__all__ = ["Model", "ModelParser", "YOLOModel"]
