"""API-compatible stand-in for the third-party `kindle` package (not vendored by the reference)."""
from . import modules
from .model import Model, ModelParser, YOLOModel

__all__ = ["Model", "ModelParser", "YOLOModel", "modules"]
