"""ctypes binding of the CPU oracle (oracle/footile_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
``footile_b200`` never imports this module.

The oracle is a C++ restatement of footile's CPU rasteriser (the Rust crate
cannot be built here); see the header of footile_oracle.cpp for what pins it
to the reference and what is parity-unpinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfootile_oracle.so")

# Same 28-byte layout as include/footile_b200.h: ftl_path_op
OP_DTYPE = np.dtype([("tag", "<u4"), ("v", "<f4", (6,))])
CLOSE, MOVE, LINE, QUAD, CUBIC, PENWIDTH = range(6)
MATTE8, GRAYA8P, RGBA8P = range(3)
NONZERO, EVENODD = 0, 1
MITER, BEVEL, ROUND = range(3)
BPP = {MATTE8: 1, GRAYA8P: 2, RGBA8P: 4}


def build(force=False):
    """Compile the oracle shared library with oracle/Makefile."""
    src = os.path.join(_HERE, "footile_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfootile_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, i16p, i32p, u32p, f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_int16, C.c_int32, C.c_uint32, C.c_float))
        L.orc_fx_from_f32.restype = C.c_int32
        L.orc_fx_from_f32.argtypes = [C.c_float]
        L.orc_fx_from_i32.restype = C.c_int32
        L.orc_fx_from_i32.argtypes = [C.c_int32]
        L.orc_fx_to_f32.restype = C.c_float
        L.orc_fx_to_f32.argtypes = [C.c_int32]
        L.orc_fx_to_i32.restype = C.c_int32
        L.orc_fx_to_i32.argtypes = [C.c_int32]
        L.orc_fx_op.restype = C.c_int32
        L.orc_fx_op.argtypes = [C.c_int, C.c_int32, C.c_int32]
        L.orc_widdershins.restype = C.c_int
        L.orc_widdershins.argtypes = [C.c_int32] * 4
        L.orc_pixel_cov.restype = C.c_int
        L.orc_pixel_cov.argtypes = [C.c_int32]
        L.orc_accumulate.restype = None
        L.orc_accumulate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_has_ssse3.restype = C.c_int
        L.orc_src_over.restype = None
        L.orc_src_over.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint8]
        L.orc_convert_srgb.restype = None
        L.orc_convert_srgb.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_fig_fill.restype = C.c_int
        L.orc_fig_fill.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_plotter_new.restype = C.c_void_p
        L.orc_plotter_new.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.orc_plotter_free.restype = None
        L.orc_plotter_free.argtypes = [C.c_void_p]
        L.orc_set_tolerance.restype = None
        L.orc_set_tolerance.argtypes = [C.c_void_p, C.c_float]
        L.orc_set_transform.restype = None
        L.orc_set_transform.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_join.restype = None
        L.orc_set_join.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.orc_set_options.restype = None
        L.orc_set_options.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int]
        L.orc_fill.restype = C.c_int
        L.orc_fill.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_stroke.restype = C.c_int
        L.orc_stroke.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_read_raster.restype = None
        L.orc_read_raster.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_write_raster.restype = None
        L.orc_write_raster.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_last_info.restype = None
        L.orc_last_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_rows.restype = None
        L.orc_set_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.orc_batch_fill_timed.restype = C.c_double
        L.orc_batch_fill_timed.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_batch_fill_checksums.restype = None
        L.orc_batch_fill_checksums.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_debug_edges.restype = C.c_size_t
        L.orc_debug_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_get_pen_width.restype = C.c_float
        L.orc_get_pen_width.argtypes = [C.c_void_p]
        L.orc_debug_flatten.restype = C.c_size_t
        L.orc_debug_flatten.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                        C.c_size_t, C.c_void_p]
        L.orc_debug_flatten_wide.restype = C.c_size_t
        L.orc_debug_flatten_wide.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_debug_stroke_ops.restype = C.c_size_t
        L.orc_debug_stroke_ops.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def _ops(ops):
    a = np.ascontiguousarray(np.asarray(ops, dtype=OP_DTYPE))
    return a, a.ctypes.data, len(a)


def _clr(clr, fmt):
    c = np.zeros(4, dtype=np.uint8)
    if clr is not None:
        v = np.asarray(clr, dtype=np.uint8).ravel()
        c[: len(v)] = v
    return c


class Fixed:
    """Fixed 16.16 helpers (src/fixed.rs) for the KAT tests."""
    OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "shl": 4, "shr": 5, "abs": 6, "floor": 7, "ceil": 8, "round": 9,
           "trunc": 10, "fract": 11, "avg": 12}

    @staticmethod
    def f(x):
        return lib().orc_fx_from_f32(float(x))

    @staticmethod
    def i(x):
        return lib().orc_fx_from_i32(int(x))

    @staticmethod
    def op(name, a, b=0):
        return lib().orc_fx_op(Fixed.OPS[name], a, b)

    @staticmethod
    def to_f32(a):
        return lib().orc_fx_to_f32(a)

    @staticmethod
    def to_i32(a):
        return lib().orc_fx_to_i32(a)


def accumulate(rule, src, simd=True):
    """imgbuf.rs accumulate_non_zero / accumulate_even_odd. Returns (dst u8, zeroed src)."""
    s = np.array(src, dtype=np.int16, copy=True)
    d = np.zeros(len(s), dtype=np.uint8)
    lib().orc_accumulate(rule, d.ctypes.data, s.ctypes.data, len(s), 1 if simd else 0)
    return d, s


def src_over(dst, src, alpha):
    d = np.array(dst, dtype=np.uint8, copy=True)
    s = np.asarray(src, dtype=np.uint8)
    lib().orc_src_over(d.ctypes.data, s.ctypes.data, len(d), int(alpha))
    return d


def fig_fill(w, h, fmt, rule, subs, clr=None, raster=None, mode=0, simd=True, vid_cap=65535, want_area=False):
    """Fig::add_point.. / close / fill (src/fig.rs) on explicit sub-figures.

    subs: list of lists of (x, y) floats.  Returns (raster[h, w*bpp] u8, info dict[, area i16[h,w]]).
    """
    pts = np.array([p for s in subs for p in s], dtype=np.float32).reshape(-1, 2)
    lens = np.array([len(s) for s in subs], dtype=np.uint32)
    bpp = BPP[fmt]
    ras = np.zeros((h, w * bpp), dtype=np.uint8) if raster is None else np.array(raster, dtype=np.uint8, copy=True).reshape(h, w * bpp)
    area = np.zeros((h, w), dtype=np.int16) if want_area else None
    info = np.zeros(3, dtype=np.int32)
    c = _clr(clr, fmt)
    lib().orc_fig_fill(w, h, fmt, rule, pts.ctypes.data, lens.ctypes.data, len(lens), c.ctypes.data, ras.ctypes.data,
                       mode, 1 if simd else 0, vid_cap, area.ctypes.data if want_area else None, info.ctypes.data)
    d = {"dir": int(info[0]), "top_row": int(info[1]), "n_points": int(info[2])}
    return (ras, d, area) if want_area else (ras, d)


def batch_fill_timed(w, h, fmt, ops, offs, rules=None, transforms=None, clr=(255, 255, 255, 255), threads=1, repeats=1):
    """Seconds the oracle needs to fill job j = ops[offs[j]:offs[j+1]] into its own w x h raster, for all jobs,
    `repeats` times, on `threads` C++ threads (rasters pre-allocated, timed inside C++)."""
    a = np.ascontiguousarray(np.asarray(ops, dtype=OP_DTYPE))
    o = np.ascontiguousarray(np.asarray(offs, dtype=np.uint64))
    r = None if rules is None else np.ascontiguousarray(np.asarray(rules, dtype=np.uint8))
    t = None if transforms is None else np.ascontiguousarray(np.asarray(transforms, dtype=np.float32))
    c = _clr(clr, fmt)
    return lib().orc_batch_fill_timed(w, h, fmt, len(o) - 1, a.ctypes.data, o.ctypes.data, None if r is None else r.ctypes.data,
                                      None if t is None else t.ctypes.data, c.ctypes.data, int(threads), int(repeats))


def batch_fill_checksums(w, h, fmt, ops, offs, rules=None, transforms=None, clr=(255, 255, 255, 255), threads=1):
    """FNV checksum (the fold of ftl_batch_checksums) of every job's raster, computed by the oracle on `threads` threads."""
    a = np.ascontiguousarray(np.asarray(ops, dtype=OP_DTYPE))
    o = np.ascontiguousarray(np.asarray(offs, dtype=np.uint64))
    r = None if rules is None else np.ascontiguousarray(np.asarray(rules, dtype=np.uint8))
    t = None if transforms is None else np.ascontiguousarray(np.asarray(transforms, dtype=np.float32))
    c = _clr(clr, fmt)
    out = np.zeros(len(o) - 1, dtype=np.uint64)
    lib().orc_batch_fill_checksums(w, h, fmt, len(o) - 1, a.ctypes.data, o.ctypes.data, None if r is None else r.ctypes.data,
                                   None if t is None else t.ctypes.data, c.ctypes.data, int(threads), out.ctypes.data)
    return out


class Plotter:
    """Mirror of footile::Plotter (src/plotter.rs:38-380) over the oracle."""

    def __init__(self, width, height, fmt=MATTE8, init=None, vid_cap=65535, simd=True, orderfree=False):
        self.width, self.height, self.fmt = width, height, fmt
        self.bpp = BPP[fmt]
        buf = None
        if init is not None:
            buf = np.ascontiguousarray(np.asarray(init, dtype=np.uint8)).ravel()
            assert buf.size == width * height * self.bpp
        self._h = lib().orc_plotter_new(width, height, fmt, buf.ctypes.data if buf is not None else None)
        lib().orc_set_options(self._h, vid_cap, 1 if simd else 0, 1 if orderfree else 0)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_plotter_free(self._h)
            self._h = None

    def set_tolerance(self, t):
        lib().orc_set_tolerance(self._h, float(t))
        return self

    def set_transform(self, e):
        a = np.asarray(e, dtype=np.float32)
        assert a.size == 6
        lib().orc_set_transform(self._h, a.ctypes.data)
        return self

    def set_join(self, kind, miter_limit=4.0):
        lib().orc_set_join(self._h, kind, float(miter_limit))
        return self

    def set_rows(self, lo, hi):
        """Order-free mode only: draw raster rows [lo, hi) only (stripe checks of huge rasters)."""
        lib().orc_set_rows(self._h, int(lo), int(hi))
        return self

    def fill(self, rule, ops, clr=None):
        a, p, n = _ops(ops)
        c = _clr(clr, self.fmt)
        lib().orc_fill(self._h, rule, p, n, c.ctypes.data)
        return self

    def stroke(self, ops, clr=None):
        a, p, n = _ops(ops)
        c = _clr(clr, self.fmt)
        lib().orc_stroke(self._h, p, n, c.ctypes.data)
        return self

    def raster(self):
        out = np.zeros((self.height, self.width * self.bpp), dtype=np.uint8)
        lib().orc_read_raster(self._h, out.ctypes.data)
        return out

    def write_raster(self, px):
        buf = np.ascontiguousarray(np.asarray(px, dtype=np.uint8)).ravel()
        assert buf.size == self.width * self.height * self.bpp
        lib().orc_write_raster(self._h, buf.ctypes.data)

    def last_info(self):
        info = np.zeros(3, dtype=np.int32)
        lib().orc_last_info(self._h, info.ctypes.data)
        return {"dir": int(info[0]), "top_row": int(info[1]), "n_points": int(info[2])}

    def pen_width(self):
        return lib().orc_get_pen_width(self._h)

    def debug_flatten(self, ops):
        a, p, n = _ops(ops)
        cap = 1 << 16
        while True:
            xy = np.zeros((cap, 2), dtype=np.int32)
            subs = np.zeros((cap, 2), dtype=np.uint32)
            ns = C.c_size_t(0)
            npts = lib().orc_debug_flatten(self._h, p, n, xy.ctypes.data, cap, subs.ctypes.data, cap, C.byref(ns))
            if npts <= cap and ns.value <= cap:
                return xy[:npts].copy(), subs[: ns.value].copy()
            cap = max(npts, ns.value)

    def debug_edges(self, ops):
        """(n, 6) int32: x_bot, inv_slope, step_pix, y_upper, y_lower, sign of every edge Fig::fill builds (fig.rs:179-210,286)."""
        a, p, n_ops = _ops(ops)
        cap = 1 << 14
        while True:
            out = np.zeros((cap, 6), dtype=np.int32)
            n = lib().orc_debug_edges(self._h, p, n_ops, out.ctypes.data, cap)
            if n <= cap:
                return out[:n]
            cap = n

    def debug_flatten_wide(self, ops):
        a, p, n = _ops(ops)
        cap = 1 << 16
        while True:
            out = np.zeros((cap, 3), dtype=np.float32)
            k = lib().orc_debug_flatten_wide(self._h, p, n, out.ctypes.data, cap)
            if k <= cap:
                return out[:k].copy()
            cap = k

    def debug_stroke_ops(self, ops):
        a, p, n = _ops(ops)
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=OP_DTYPE)
            k = lib().orc_debug_stroke_ops(self._h, p, n, out.ctypes.data, cap)
            if k <= cap:
                return out[:k].copy()
            cap = k


def convert_srgb(fmt, pixels):
    """Output conversion of examples/fishy.rs:33 on the CPU: Rgba8p -> SRgba8 / Graya8p -> SGraya8 / Matte8 -> SGray8 bytes."""
    a = np.ascontiguousarray(np.asarray(pixels, dtype=np.uint8))
    out = np.empty_like(a)
    lib().orc_convert_srgb(int(fmt), a.ctypes.data, out.ctypes.data, a.size // BPP[fmt])
    return out
