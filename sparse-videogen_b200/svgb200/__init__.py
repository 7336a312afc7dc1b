"""svgb200 — B200-native sparse video-DiT attention engine (host side).

Python mirror of the reference's operator interface (svg/kmeans_utils.py, svg/kernels/triton/permute.py,
svg/models/*/placement.py, svg/kernels/ops) on top of the C-ABI library libsvgb200.so.
"""
from . import core  # noqa: F401
from ._lib import SvgbError, lib, lib_path  # noqa: F401

__version__ = "0.1.0"
