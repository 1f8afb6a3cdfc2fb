"""Plotter: host-side mirror of footile::Plotter over the C ABI.

Same names, argument meaning and behaviour as the reference
(src/plotter.rs:38-380); the raster lives in HBM and every drawing call runs
the CUDA pipeline.  Pixel formats follow pix: ``Matte8`` (1 byte), ``Graya8p``
(gray, alpha) and ``Rgba8p`` (r, g, b, a), premultiplied, row-major.
"""
import ctypes as C
import enum

import numpy as np

from . import _lib
from .path import FillRule, JoinStyle, OP_DTYPE, as_ops


class Format(enum.IntEnum):
    Matte8 = 0
    Graya8p = 1
    Rgba8p = 2


BPP = {Format.Matte8: 1, Format.Graya8p: 2, Format.Rgba8p: 4}


class Raster:
    """pix::Raster<P> on the host: ``pixels`` is a (height, width*bpp) u8 array."""

    def __init__(self, width, height, fmt=Format.Matte8, pixels=None):
        self.width, self.height, self.fmt = int(width), int(height), Format(fmt)
        bpp = BPP[self.fmt]
        if pixels is None:
            self.pixels = np.zeros((self.height, self.width * bpp), dtype=np.uint8)
        else:
            self.pixels = np.ascontiguousarray(np.asarray(pixels, dtype=np.uint8)).reshape(self.height, self.width * bpp)

    @classmethod
    def with_clear(cls, width, height, fmt=Format.Matte8):
        return cls(width, height, fmt)

    @classmethod
    def with_color(cls, width, height, fmt, clr):
        r = cls(width, height, fmt)
        bpp = BPP[r.fmt]
        r.pixels.reshape(height, width, bpp)[:] = np.asarray(clr, dtype=np.uint8)[:bpp]
        return r

    def as_u8_slice(self):
        return self.pixels.ravel()


def _color(clr, bpp):
    c = np.zeros(4, dtype=np.uint8)
    if clr is None:
        c[:] = 255
    else:
        v = np.asarray(clr, dtype=np.uint8).ravel()
        c[: min(len(v), 4)] = v[:4]
    return c


class Plotter:
    """Plotter for 2D vector paths (src/plotter.rs:38-56).

    ``Plotter(raster)`` takes ownership of the raster's pixels (copied to the
    device); ``raster()`` reads them back.  ``rows=(begin, end)`` makes this
    plotter own only a row band of the raster (multi-GPU split of one raster).
    """

    def __init__(self, raster, device=0, rows=None):
        self._setup(raster.width, raster.height, raster.fmt, device, rows, raster.pixels)

    @classmethod
    def with_clear(cls, width, height, fmt=Format.Matte8, device=0, rows=None):
        """Plotter over a cleared raster (Raster::with_clear) without materialising it on the host."""
        self = cls.__new__(cls)
        self._setup(width, height, Format(fmt), device, rows, None)
        return self

    def _setup(self, width, height, fmt, device, rows, pixels):
        self.fmt = fmt
        self._w, self._h = int(width), int(height)
        self._rows = (0, self._h) if rows is None else (int(rows[0]), int(rows[1]))
        self._bpp = BPP[self.fmt]
        self._handle = C.c_void_p()
        band = None if pixels is None else np.ascontiguousarray(pixels[self._rows[0]: self._rows[1]])
        _lib.check(_lib.lib().ftl_plotter_new_band(self._w, self._h, self._rows[0], self._rows[1], int(self.fmt),
                                                   band.ctypes.data if band is not None and band.size else None, device,
                                                   C.byref(self._handle)))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            _lib.lib().ftl_plotter_free(h)
            self._handle = None

    def width(self):
        return int(_lib.lib().ftl_width(self._handle))

    def height(self):
        return int(_lib.lib().ftl_height(self._handle))

    def set_tolerance(self, t):
        _lib.check(_lib.lib().ftl_set_tolerance(self._handle, float(t)))
        return self

    def set_transform(self, t):
        """t: pointy Transform as 6 floats [a, b, tx, c, d, ty]: x' = a*x + b*y + tx, y' = c*x + d*y + ty."""
        e = np.ascontiguousarray(np.asarray(t, dtype=np.float32).ravel())
        if e.size != 6:
            raise ValueError("transform needs 6 floats")
        _lib.check(_lib.lib().ftl_set_transform(self._handle, e.ctypes.data))
        return self

    def set_join(self, js):
        _lib.check(_lib.lib().ftl_set_join(self._handle, js.kind, js.limit))
        return self

    def set_strict_vid(self, on=True):
        """Reproduce the reference's Vid(u16) cap: Fig::add_point ignores points while 65 535 are stored (fig.rs:430)."""
        _lib.check(_lib.lib().ftl_set_strict_vid(self._handle, 1 if on else 0))
        return self

    def pen_width(self):
        return float(_lib.lib().ftl_pen_width(self._handle))

    def fill(self, rule, ops, clr=None):
        a = as_ops(ops)
        c = _color(clr, self._bpp)
        _lib.check(_lib.lib().ftl_fill(self._handle, int(rule), a.ctypes.data if len(a) else None, len(a), c.ctypes.data))
        return self

    def upload(self, rule, ops, clr=None):
        """Make a fill resident in HBM without drawing it; replay() then draws it with no host traffic."""
        a = as_ops(ops)
        c = _color(clr, self._bpp)
        _lib.check(_lib.lib().ftl_fill_upload(self._handle, int(rule), a.ctypes.data if len(a) else None, len(a), c.ctypes.data))
        return self

    def replay(self):
        _lib.check(_lib.lib().ftl_fill_replay(self._handle))
        return self

    def time_fills(self, rule, ops, clr=None, iters=1000, sync_each=True):
        """Microseconds per fill call, timed inside the library (no ctypes / numpy overhead in the figure)."""
        a = as_ops(ops)
        c = _color(clr, self._bpp)
        us = C.c_double(0)
        _lib.check(_lib.lib().ftl_time_fills(self._handle, int(rule), a.ctypes.data if len(a) else None, len(a), c.ctypes.data, int(iters),
                                             1 if sync_each else 0, C.byref(us)))
        return us.value

    def stroke(self, ops, clr=None):
        a = as_ops(ops)
        c = _color(clr, self._bpp)
        _lib.check(_lib.lib().ftl_stroke(self._handle, a.ctypes.data if len(a) else None, len(a), c.ctypes.data))
        return self

    def stroke_outline(self, ops):
        """The outline ops `stroke` would fill (updates the persistent pen width like `stroke`)."""
        a = as_ops(ops)
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=OP_DTYPE)
            n = C.c_size_t()
            _lib.check(_lib.lib().ftl_stroke_outline(self._handle, a.ctypes.data if len(a) else None, len(a), out.ctypes.data, cap, C.byref(n)))
            if n.value <= cap:
                return out[: n.value].copy()
            cap = n.value

    def fill_layers(self, layers):
        """Draw a scene in one pass: layers = [(rule, ops, color), ...] composited in order, equal to
        calling fill() once per layer.  Use stroke_outline() to turn a stroke into a NonZero layer."""
        arrs = [as_ops(o) for _, o, _ in layers]
        offs = np.zeros(len(arrs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(a) for a in arrs])
        ops = np.ascontiguousarray(np.concatenate(arrs)) if arrs else np.zeros(0, dtype=OP_DTYPE)
        rules = np.array([int(r) for r, _, _ in layers], dtype=np.uint8)
        colors = np.ascontiguousarray(np.array([_color(c, self._bpp) for _, _, c in layers], dtype=np.uint8).reshape(-1, 4))
        _lib.check(_lib.lib().ftl_fill_layers(self._handle, len(layers), ops.ctypes.data if len(ops) else None, offs.ctypes.data,
                                              rules.ctypes.data if len(layers) else None, colors.ctypes.data if len(layers) else None))
        return self

    def stream(self):
        """The cudaStream_t (as int) every call on this plotter is issued on."""
        st = C.c_void_p()
        _lib.check(_lib.lib().ftl_stream(self._handle, C.byref(st)))
        return st.value or 0

    def sync(self):
        _lib.check(_lib.lib().ftl_sync(self._handle))
        return self

    def tile_kernel_time(self, reset=False):
        """(ms, launches) of this handle's tile kernels while profiling is on (the state is per handle)."""
        ms, n = C.c_double(0), C.c_uint64(0)
        _lib.check(_lib.lib().ftl_plotter_tile_kernel_time(self._handle, 1 if reset else 0, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def raster(self):
        """Read the owned rows back: Raster of (rows, width)."""
        n_rows = self._rows[1] - self._rows[0]
        out = np.empty((n_rows, self._w * self._bpp), dtype=np.uint8)
        _lib.check(_lib.lib().ftl_read_raster(self._handle, out.ctypes.data if out.size else None, out.size))
        return Raster(self._w, n_rows, self.fmt, out)

    def into_raster(self):
        return self.raster()

    def raster_srgb(self):
        """The owned rows converted for output on the device (examples/fishy.rs:33: SRgba8 / SGraya8 / SGray8 bytes)."""
        n_rows = self._rows[1] - self._rows[0]
        out = np.empty((n_rows, self._w * self._bpp), dtype=np.uint8)
        _lib.check(_lib.lib().ftl_read_raster_srgb(self._handle, out.ctypes.data if out.size else None, out.size))
        return out

    def write_raster(self, pixels):
        a = np.ascontiguousarray(np.asarray(pixels, dtype=np.uint8)).ravel()
        _lib.check(_lib.lib().ftl_write_raster(self._handle, a.ctypes.data if a.size else None, a.size))

    def device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        _lib.check(_lib.lib().ftl_raster_device_ptr(self._handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    # ---- parity probes ----
    def debug_flatten(self, ops):
        a = as_ops(ops)
        cap = 1 << 16
        while True:
            xy = np.zeros((cap, 2), dtype=np.int32)
            subs = np.zeros((cap, 2), dtype=np.uint32)
            npts, nsub = C.c_size_t(), C.c_size_t()
            _lib.check(_lib.lib().ftl_debug_flatten(self._handle, a.ctypes.data if len(a) else None, len(a), xy.ctypes.data, cap,
                                                    C.byref(npts), subs.ctypes.data, cap, C.byref(nsub)))
            if npts.value <= cap and nsub.value <= cap:
                return xy[: npts.value].copy(), subs[: nsub.value].copy()
            cap = max(npts.value, nsub.value)

    def debug_last_fill(self):
        info = np.zeros(3, dtype=np.int32)
        _lib.check(_lib.lib().ftl_debug_last_fill(self._handle, info.ctypes.data))
        return {"dir": int(info[0]), "top_row": int(info[1]), "n_points": int(info[2])}

    def debug_area(self, row):
        """int16[width]: the signed-area deltas of raster row `row` of the last fill before the prefix sum (stage (c) probe)."""
        out = np.zeros(self._w, dtype=np.int16)
        _lib.check(_lib.lib().ftl_debug_area(self._handle, int(row), out.ctypes.data, self._w))
        return out

    def debug_edges(self):
        """(n, 6) int32 edges of the last fill: x_bot, inv_slope, step_pix, y_upper, y_lower, sign (stage (b) probe)."""
        n = C.c_size_t(0)
        cap = 1024
        while True:
            out = np.zeros((cap, 6), dtype=np.int32)
            _lib.check(_lib.lib().ftl_debug_edges(self._handle, out.ctypes.data, cap, C.byref(n)))
            if n.value <= cap:
                return out[:n.value]
            cap = n.value

    def debug_stroke_ops(self, ops):
        a = as_ops(ops)
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=OP_DTYPE)
            n = C.c_size_t()
            _lib.check(_lib.lib().ftl_debug_stroke_ops(self._handle, a.ctypes.data if len(a) else None, len(a), out.ctypes.data, cap,
                                                       C.byref(n)))
            if n.value <= cap:
                return out[: n.value].copy()
            cap = n.value

    def debug_stroke_ops_device(self, ops):
        """The outline as the device stroker builds it (stroke_kernels.cuh), or None when it declined the call."""
        a = as_ops(ops)
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=OP_DTYPE)
            n, fb = C.c_size_t(), C.c_int()
            _lib.check(_lib.lib().ftl_debug_stroke_ops_device(self._handle, a.ctypes.data if len(a) else None, len(a), out.ctypes.data, cap,
                                                              C.byref(n), C.byref(fb)))
            if fb.value:
                return None
            if n.value <= cap:
                return out[: n.value].copy()
            cap = n.value


def debug_stroke_subs(ops):
    """(n, 4) uint32: first drawing op, one past the last, joined, 0 for every sub-stroke of the path (stroker.rs:204-236). Pure host code."""
    a = as_ops(ops)
    out = np.zeros((max(len(a), 1), 4), dtype=np.uint32)
    n = C.c_size_t()
    _lib.check(_lib.lib().ftl_debug_stroke_subs(a.ctypes.data if len(a) else None, len(a), out.ctypes.data, len(out), C.byref(n)))
    return out[: n.value].copy()


def debug_strict_intake(ops, transform=(1, 0, 0, 0, 1, 0), tolerance=0.3):
    """The host-side point intake of strict Vid(u16) mode: the Move / Line ops a capped fill hands to the device, or None when the
    fill stays below the 65 535-point cap. Pure host code."""
    a = as_ops(ops)
    e = np.ascontiguousarray(np.asarray(transform, dtype=np.float32))
    cap = 1 << 17
    out = np.zeros(cap, dtype=OP_DTYPE)
    n, capped = C.c_size_t(), C.c_int()
    _lib.check(_lib.lib().ftl_debug_strict_intake(e.ctypes.data, float(tolerance), a.ctypes.data if len(a) else None, len(a), out.ctypes.data, cap,
                                                  C.byref(n), C.byref(capped)))
    return out[: n.value].copy() if capped.value else None


def debug_libm_selftest(n, seed=0):
    """libm_compat.cuh against this host's libm: (hypotf mismatches, atan2f mismatches, wrong |sin| predictions, undecided
    |sin| predictions) over 4 * n random inputs and every float next to +-pi/2. Pure host code."""
    h, a, sm, su = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    _lib.check(_lib.lib().ftl_debug_libm_selftest(int(n), int(seed), C.byref(h), C.byref(a), C.byref(sm), C.byref(su)))
    return h.value, a.value, sm.value, su.value


def debug_accumulate(rule, src, device=0):
    """Row accumulate alone (imgbuf.rs:38-51,141-154) on the device. src: (rows, n) or (n,) i16."""
    s = np.ascontiguousarray(np.asarray(src, dtype=np.int16))
    rows = 1 if s.ndim == 1 else s.shape[0]
    n = s.shape[-1]
    dst = np.zeros(s.shape, dtype=np.uint8)
    _lib.check(_lib.lib().ftl_debug_accumulate(int(rule), s.ctypes.data, dst.ctypes.data, n, rows, device))
    return dst


class Batch:
    """Many independent fills per launch: ``capacity`` rasters of one size/format in HBM."""

    def __init__(self, width, height, fmt=Format.Matte8, capacity=1, device=0):
        self.width, self.height, self.fmt, self.capacity = int(width), int(height), Format(fmt), int(capacity)
        self._bpp = BPP[self.fmt]
        self._handle = C.c_void_p()
        _lib.check(_lib.lib().ftl_batch_new(self.width, self.height, int(self.fmt), self.capacity, device, C.byref(self._handle)))
        self._keep = None

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            _lib.lib().ftl_batch_free(h)
            self._handle = None

    def set_tolerance(self, t):
        _lib.check(_lib.lib().ftl_batch_set_tolerance(self._handle, float(t)))
        return self

    def clear(self, first=0, count=None):
        _lib.check(_lib.lib().ftl_batch_clear(self._handle, first, self.capacity - first if count is None else count))
        return self

    @staticmethod
    def pack(paths):
        """Concatenate a list of op arrays: returns (ops, offsets u64[n+1])."""
        arrs = [as_ops(p) for p in paths]
        offs = np.zeros(len(arrs) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(a) for a in arrs])
        ops = np.concatenate(arrs) if arrs else np.zeros(0, dtype=OP_DTYPE)
        return np.ascontiguousarray(ops), offs

    def _args(self, ops, offsets, rules, transforms, colors):
        ops = np.ascontiguousarray(np.asarray(ops, dtype=OP_DTYPE))
        offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.uint64))
        n = len(offsets) - 1
        r = None if rules is None else np.ascontiguousarray(np.asarray(rules, dtype=np.uint8))
        t = None if transforms is None else np.ascontiguousarray(np.asarray(transforms, dtype=np.float32).reshape(n, 6))
        c = None if colors is None else np.ascontiguousarray(np.asarray(colors, dtype=np.uint8).reshape(n, 4))
        self._keep = (ops, offsets, r, t, c)
        p = lambda a: None if a is None or a.size == 0 else a.ctypes.data
        return n, p(ops), offsets.ctypes.data, p(r), p(t), p(c)

    def fill(self, ops, offsets, rules=None, transforms=None, colors=None):
        _lib.check(_lib.lib().ftl_batch_fill(self._handle, *self._args(ops, offsets, rules, transforms, colors)))
        return self

    def upload(self, ops, offsets, rules=None, transforms=None, colors=None):
        _lib.check(_lib.lib().ftl_batch_upload(self._handle, *self._args(ops, offsets, rules, transforms, colors)))
        return self

    def set_join(self, js):
        _lib.check(_lib.lib().ftl_batch_set_join(self._handle, js.kind, js.limit))
        return self

    def stroke(self, ops, offsets, transforms=None, colors=None):
        """Stroke path j into raster j (every job like a new Plotter: pen width 1, this batch's join and tolerance)."""
        n, po, pf, _, pt, pc = self._args(ops, offsets, None, transforms, colors)
        _lib.check(_lib.lib().ftl_batch_stroke(self._handle, n, po, pf, pt, pc))
        return self

    def run(self):
        _lib.check(_lib.lib().ftl_batch_run(self._handle))
        return self

    def sync(self):
        _lib.check(_lib.lib().ftl_batch_sync(self._handle))
        return self

    def tile_kernel_time(self, reset=False):
        """(ms, launches) of this handle's tile kernels while profiling is on (the state is per handle)."""
        ms, n = C.c_double(0), C.c_uint64(0)
        _lib.check(_lib.lib().ftl_batch_tile_kernel_time(self._handle, 1 if reset else 0, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def read(self, first=0, count=None, out=None):
        count = self.capacity - first if count is None else count
        shape = (count, self.height, self.width * self._bpp)
        if out is None:
            out = np.empty(shape, dtype=np.uint8)
        _lib.check(_lib.lib().ftl_batch_read(self._handle, first, count, out.ctypes.data if out.size else None, out.size))
        return out.reshape(shape)

    def checksums(self, first=0, count=None):
        count = self.capacity - first if count is None else count
        out = np.zeros(count, dtype=np.uint64)
        _lib.check(_lib.lib().ftl_batch_checksums(self._handle, first, count, out.ctypes.data))
        return out

    def top_rows(self, first=0, count=None):
        """top_row of each job of the last fill / run (int32; INT32_MAX for a job that drew nothing)."""
        count = self.capacity - first if count is None else count
        out = np.zeros(count, dtype=np.int32)
        _lib.check(_lib.lib().ftl_batch_debug_top_rows(self._handle, first, count, out.ctypes.data))
        return out

    def device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        _lib.check(_lib.lib().ftl_batch_device_ptr(self._handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stream(self):
        """cudaStream_t (as int) the batch issues its work on."""
        s = C.c_void_p()
        _lib.check(_lib.lib().ftl_batch_stream(self._handle, C.byref(s)))
        return s.value

    def read_into(self, first, count, ptr, nbytes):
        """Copy rasters to a caller-provided host address (e.g. pinned memory)."""
        _lib.check(_lib.lib().ftl_batch_read(self._handle, first, count, C.c_void_p(ptr), nbytes))
