"""footile_b200 — B200-native 2D path rasteriser with footile's Plotter/Path2D API.

The hot path (curve flattening -> fixed-point edges -> signed-area coverage
scatter -> row prefix sum + fill rule + store/blend) runs as hand-written
sm_100a CUDA kernels in ``libfootile_b200.so`` behind the C ABI declared in
``include/footile_b200.h``.  There is no CPU fallback: without the shared
library or a CUDA device every drawing call raises.
"""
from .path import FillRule, JoinStyle, OpTag, OP_DTYPE, Path2D, PathOp, as_ops  # noqa: F401
from .plotter import Batch, Format, Plotter, Raster, debug_accumulate  # noqa: F401
from ._lib import FootileError, device_count, launch_count, set_profiling, tile_kernel_time, transfer_bytes  # noqa: F401

__all__ = ["FillRule", "JoinStyle", "OpTag", "OP_DTYPE", "Path2D", "PathOp", "as_ops", "Batch", "Format", "Plotter", "Raster",
           "debug_accumulate", "FootileError", "device_count", "launch_count", "set_profiling", "tile_kernel_time", "transfer_bytes"]
