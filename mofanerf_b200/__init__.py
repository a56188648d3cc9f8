"""mofanerf_b200 — B200-native (sm_100a) volume-rendering engine for MoFaNeRF's ray-marching hot path.

Public surface:
  B200Renderer   drop-in for the reference's models.render_class.myRenderer
  install()      swap it into an importable reference tree
  Engine         ctypes handle on libmofa_b200.so (include/mofa_b200.h)
  nets           parameter containers / host-side PyTorch modules (StyleModule, TexEncoder)
"""
from . import nets  # noqa: F401
from .engine import Engine, get_engine  # noqa: F401
from .renderer import B200Renderer, install, wait_for_images  # noqa: F401

__all__ = ["B200Renderer", "install", "wait_for_images", "Engine", "get_engine", "nets"]
