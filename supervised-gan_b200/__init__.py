"""supervised-gan hot path, B200-native: drop-in `networks` (define_G / define_D / GANLoss / WeightedL1Loss)
and step drivers backed by hand-written sm_100a kernels (libsgk.so, C ABI in include/sgk.h)."""
from . import _lib, ops, networks  # noqa: F401
from .ops import set_precision, get_precision  # noqa: F401

__all__ = ["networks", "ops", "set_precision", "get_precision"]
