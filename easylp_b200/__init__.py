"""easylp_b200 — B200-native solve path behind EasyLP's R6 API (Python host mirror + ctypes ABI binding).

The CUDA library (libeasylp_b200.so) is the product; there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
