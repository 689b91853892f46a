"""kindle.modules — operator classes named in the model yamls."""
from .activation import Activation
from .bottleneck import C3, Bottleneck, BottleneckCSP
from .concat import Concat
from .conv import Conv, Focus
from .poolings import SPP, SPPF
from .yolo_head import YOLOHead

__all__ = ["Activation", "Bottleneck", "BottleneckCSP", "C3", "Concat", "Conv", "Focus", "SPP", "SPPF", "YOLOHead"]
