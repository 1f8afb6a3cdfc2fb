"""ctypes loader of libfootile_b200.so (the C ABI of include/footile_b200.h).

The shared library is the product: if it is missing, or there is no CUDA
device when a compute entry point is called, this module raises — there is no
CPU fallback anywhere in the package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfootile_b200.so")

# every symbol include/footile_b200.h declares (tests check the .so exports them all)
SYMBOLS = [
    "ftl_abi_version", "ftl_last_error", "ftl_device_count",
    "ftl_plotter_new", "ftl_plotter_new_band", "ftl_plotter_free", "ftl_width", "ftl_height",
    "ftl_set_tolerance", "ftl_set_transform", "ftl_set_join", "ftl_set_strict_vid", "ftl_pen_width",
    "ftl_fill", "ftl_stroke", "ftl_fill_layers", "ftl_stroke_outline", "ftl_read_raster", "ftl_read_raster_srgb", "ftl_write_raster", "ftl_sync", "ftl_raster_device_ptr",
    "ftl_batch_new", "ftl_batch_free", "ftl_batch_set_tolerance", "ftl_batch_clear", "ftl_batch_fill", "ftl_batch_set_join", "ftl_batch_stroke",
    "ftl_batch_read", "ftl_batch_checksums", "ftl_batch_sync", "ftl_batch_device_ptr",
    "ftl_batch_upload", "ftl_batch_run", "ftl_batch_stream", "ftl_stream", "ftl_fill_upload", "ftl_fill_replay",
    "ftl_ctx_new", "ftl_ctx_free", "ftl_ctx_size", "ftl_shard_range", "ftl_band_rows", "ftl_ctx_fill_batch", "ftl_ctx_fill_bands",
    "ftl_launch_count", "ftl_transfer_bytes", "ftl_set_profiling", "ftl_tile_kernel_time", "ftl_plotter_tile_kernel_time", "ftl_batch_tile_kernel_time", "ftl_time_fills",
    "ftl_debug_area", "ftl_debug_small_profile", "ftl_batch_debug_top_rows", "ftl_debug_flatten", "ftl_debug_last_fill", "ftl_debug_edges", "ftl_debug_stroke_ops", "ftl_debug_stroke_outline", "ftl_debug_accumulate",
    "ftl_debug_stroke_ops_device", "ftl_debug_libm_selftest", "ftl_debug_strict_intake", "ftl_debug_stroke_subs",
]


class FootileError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("footile_b200 status %d: %s" % (status, msg))
        self.status = status


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C footile_b200/csrc` (no CPU fallback exists)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, sz, u32, i32, f32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_float
    sig = {
        "ftl_abi_version": (i32, []),
        "ftl_last_error": (C.c_char_p, []),
        "ftl_device_count": (i32, [vp]),
        "ftl_plotter_new": (i32, [u32, u32, i32, vp, i32, vp]),
        "ftl_plotter_new_band": (i32, [u32, u32, u32, u32, i32, vp, i32, vp]),
        "ftl_plotter_free": (i32, [vp]),
        "ftl_width": (u32, [vp]),
        "ftl_height": (u32, [vp]),
        "ftl_set_tolerance": (i32, [vp, f32]),
        "ftl_set_transform": (i32, [vp, vp]),
        "ftl_set_join": (i32, [vp, i32, f32]),
        "ftl_set_strict_vid": (i32, [vp, i32]),
        "ftl_pen_width": (f32, [vp]),
        "ftl_fill": (i32, [vp, i32, vp, sz, vp]),
        "ftl_stroke": (i32, [vp, vp, sz, vp]),
        "ftl_fill_layers": (i32, [vp, u32, vp, vp, vp, vp]),
        "ftl_stroke_outline": (i32, [vp, vp, sz, vp, sz, vp]),
        "ftl_read_raster": (i32, [vp, vp, sz]),
        "ftl_read_raster_srgb": (i32, [vp, vp, sz]),
        "ftl_write_raster": (i32, [vp, vp, sz]),
        "ftl_sync": (i32, [vp]),
        "ftl_raster_device_ptr": (i32, [vp, vp, vp]),
        "ftl_batch_new": (i32, [u32, u32, i32, u32, i32, vp]),
        "ftl_batch_free": (i32, [vp]),
        "ftl_batch_set_tolerance": (i32, [vp, f32]),
        "ftl_batch_clear": (i32, [vp, u32, u32]),
        "ftl_batch_fill": (i32, [vp, u32, vp, vp, vp, vp, vp]),
        "ftl_batch_set_join": (i32, [vp, i32, f32]),
        "ftl_batch_stroke": (i32, [vp, u32, vp, vp, vp, vp]),
        "ftl_batch_read": (i32, [vp, u32, u32, vp, sz]),
        "ftl_batch_checksums": (i32, [vp, u32, u32, vp]),
        "ftl_batch_sync": (i32, [vp]),
        "ftl_batch_device_ptr": (i32, [vp, vp, vp]),
        "ftl_batch_upload": (i32, [vp, u32, vp, vp, vp, vp, vp]),
        "ftl_batch_run": (i32, [vp]),
        "ftl_fill_upload": (i32, [vp, i32, vp, sz, vp]),
        "ftl_fill_replay": (i32, [vp]),
        "ftl_batch_stream": (i32, [vp, vp]),
        "ftl_stream": (i32, [vp, vp]),
        "ftl_ctx_new": (i32, [i32, vp, vp]),
        "ftl_ctx_free": (i32, [vp]),
        "ftl_ctx_size": (i32, [vp]),
        "ftl_shard_range": (i32, [u32, u32, u32, vp, vp]),
        "ftl_band_rows": (i32, [u32, u32, u32, u32, vp, vp]),
        "ftl_ctx_fill_batch": (i32, [vp, u32, u32, i32, f32, u32, vp, vp, vp, vp, vp, vp, sz]),
        "ftl_ctx_fill_bands": (i32, [vp, u32, u32, i32, i32, vp, sz, vp, f32, vp, vp, vp, sz]),
        "ftl_launch_count": (C.c_uint64, []),
        "ftl_transfer_bytes": (i32, [i32, vp, vp]),
        "ftl_set_profiling": (i32, [i32]),
        "ftl_tile_kernel_time": (i32, [i32, vp, vp]),
        "ftl_plotter_tile_kernel_time": (i32, [vp, i32, vp, vp]),
        "ftl_batch_tile_kernel_time": (i32, [vp, i32, vp, vp]),
        "ftl_time_fills": (i32, [vp, i32, vp, sz, vp, u32, i32, vp]),
        "ftl_debug_flatten": (i32, [vp, vp, sz, vp, sz, vp, vp, sz, vp]),
        "ftl_debug_last_fill": (i32, [vp, vp]),
        "ftl_debug_area": (i32, [vp, i32, vp, sz]),
        "ftl_debug_small_profile": (i32, [vp, vp]),
        "ftl_batch_debug_top_rows": (i32, [vp, u32, u32, vp]),
        "ftl_debug_edges": (i32, [vp, vp, sz, vp]),
        "ftl_debug_stroke_ops": (i32, [vp, vp, sz, vp, sz, vp]),
        "ftl_debug_stroke_ops_device": (i32, [vp, vp, sz, vp, sz, vp, vp]),
        "ftl_debug_libm_selftest": (i32, [C.c_uint64, C.c_uint64, vp, vp, vp, vp]),
        "ftl_debug_strict_intake": (i32, [vp, f32, vp, sz, vp, sz, vp, vp]),
        "ftl_debug_stroke_subs": (i32, [vp, sz, vp, sz, vp]),
        "ftl_debug_stroke_outline": (i32, [i32, f32, f32, vp, sz, vp, vp, vp, sz, vp]),
        "ftl_debug_accumulate": (i32, [i32, vp, vp, sz, sz, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.ftl_abi_version() != 1:
        raise ImportError("libfootile_b200.so ABI version mismatch")
    _lib = L
    return L


def check(status):
    if status != 0:
        msg = lib().ftl_last_error()
        raise FootileError(status, msg.decode("utf-8", "replace") if msg else "")


def device_count():
    n = C.c_int(0)
    check(lib().ftl_device_count(C.byref(n)))
    return n.value


def launch_count():
    return int(lib().ftl_launch_count())


def transfer_bytes(reset=False):
    """(h2d, d2h) bytes the library moved over PCIe since load / the last reset."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    check(lib().ftl_transfer_bytes(1 if reset else 0, C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


def set_profiling(on):
    check(lib().ftl_set_profiling(1 if on else 0))


def tile_kernel_time(reset=False):
    ms = C.c_double(0)
    n = C.c_uint64(0)
    check(lib().ftl_tile_kernel_time(1 if reset else 0, C.byref(ms), C.byref(n)))
    return ms.value, int(n.value)
