"""ayolov2_b200 — B200-native (sm_100a) hot path of j-marple-dev/AYolov2.

Host side is PyTorch/Python (the reference's own language for this path) over libay2.so, a C-ABI shared
library of hand-written CUDA kernels (include/ay2.h). No CPU fallback: the ops raise if the library or a
CUDA device is missing.
"""
__version__ = "0.1.0"


def set_precision(model, precision: str = "bf16"):
    """See ayolov2_b200.engine.set_precision."""
    from .engine import set_precision as _sp

    return _sp(model, precision)
