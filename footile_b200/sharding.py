"""Multi-GPU layout of the hot path: one process per GPU, no data-path collective.

Two ways to shard:
  * independent paths / rasters: rank r fills the contiguous block ``shard_range(n, r, world)``
    of path indices into its own rasters (scene generators are counter based, so a path is the
    same whichever rank draws it);
  * one raster split into row bands: rank r owns rows ``band_rows(height, r, world)``; every rank
    receives the same ops and recomputes (dir, top_row) redundantly, so nothing is exchanged.
The only communication is the OPTIONAL final gather of bands (NCCL all-gather over NVLink on
GPUs; gloo on CPU tensors in the tests).
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous block partition of range(n): returns (first, count) of this rank (ftl_shard_range of the C ABI)."""
    import ctypes as C
    from . import _lib
    first, count = C.c_uint32(0), C.c_uint32(0)
    _lib.check(_lib.lib().ftl_shard_range(int(n), int(rank), int(world), C.byref(first), C.byref(count)))
    return int(first.value), int(count.value)


def band_rows(height, rank, world, align=32):
    """Row band [begin, end) of rank `rank`; band boundaries are multiples of `align` rows (ftl_band_rows of the C ABI;
    32 rows = one band of the binned tile kernel)."""
    import ctypes as C
    from . import _lib
    b, e = C.c_uint32(0), C.c_uint32(0)
    _lib.check(_lib.lib().ftl_band_rows(int(height), int(rank), int(world), int(align), C.byref(b), C.byref(e)))
    return int(b.value), int(e.value)


class Context:
    """Several GPUs behind one handle (ftl_ctx_* of the C ABI): one host thread per device, no data-path collective."""

    def __init__(self, devices=None):
        import ctypes as C
        from . import _lib
        self._h = C.c_void_p()
        if devices is None:
            _lib.check(_lib.lib().ftl_ctx_new(0, None, C.byref(self._h)))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            _lib.check(_lib.lib().ftl_ctx_new(len(devices), arr, C.byref(self._h)))

    def __del__(self):
        from . import _lib
        if getattr(self, "_h", None):
            _lib.lib().ftl_ctx_free(self._h)
            self._h = None

    def size(self):
        from . import _lib
        return int(_lib.lib().ftl_ctx_size(self._h))

    def fill_batch(self, width, height, fmt, ops, offsets, rules=None, transforms=None, colors=None, tolerance=0.0):
        """n independent fills sharded over the devices; returns uint8[n, height, width*bpp]."""
        from . import _lib
        from .path import OP_DTYPE
        bpp = {0: 1, 1: 2, 2: 4}[int(fmt)]
        ops = np.ascontiguousarray(np.asarray(ops, dtype=OP_DTYPE))
        offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.uint64))
        n = len(offsets) - 1
        r = None if rules is None else np.ascontiguousarray(np.asarray(rules, dtype=np.uint8))
        t = None if transforms is None else np.ascontiguousarray(np.asarray(transforms, dtype=np.float32).reshape(n, 6))
        c = None if colors is None else np.ascontiguousarray(np.asarray(colors, dtype=np.uint8).reshape(n, 4))
        out = np.empty((n, height, width * bpp), dtype=np.uint8)
        p = lambda a: None if a is None or a.size == 0 else a.ctypes.data
        _lib.check(_lib.lib().ftl_ctx_fill_batch(self._h, width, height, int(fmt), float(tolerance), n, p(ops), offsets.ctypes.data, p(r), p(t), p(c),
                                                 out.ctypes.data, out.size))
        return out

    def fill_bands(self, width, height, fmt, rule, ops, color, transform=None, tolerance=0.0, init=None):
        """One fill of one raster, one row band per device; returns uint8[height, width*bpp]."""
        from . import _lib
        from .path import OP_DTYPE
        bpp = {0: 1, 1: 2, 2: 4}[int(fmt)]
        ops = np.ascontiguousarray(np.asarray(ops, dtype=OP_DTYPE))
        clr = np.zeros(4, dtype=np.uint8)
        clr[: len(color)] = color
        tr = None if transform is None else np.ascontiguousarray(np.asarray(transform, dtype=np.float32).ravel())
        ini = None if init is None else np.ascontiguousarray(np.asarray(init, dtype=np.uint8))
        out = np.empty((height, width * bpp), dtype=np.uint8)
        _lib.check(_lib.lib().ftl_ctx_fill_bands(self._h, width, height, int(fmt), int(rule), ops.ctypes.data if len(ops) else None, len(ops),
                                                 None if tr is None else tr.ctypes.data, float(tolerance), clr.ctypes.data,
                                                 None if ini is None else ini.ctypes.data, out.ctypes.data, out.size))
        return out


def device_tensor(ptr, nbytes, device):
    """Zero-copy torch.uint8 view of `nbytes` of device memory at `ptr` (a raster owned by a handle)."""
    import torch

    class _Mem:
        __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}

    return torch.as_tensor(_Mem(), device=torch.device("cuda", device))


def gather_bands(band, height, width_bytes, group=None):
    """All-gather row bands into the full raster on every rank.

    band: torch.uint8 tensor [rows_of_this_rank, width_bytes] (CUDA with NCCL, CPU with gloo).
    Bands may differ in height by one alignment unit, so each is padded to the largest band.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    spans = [band_rows(height, r, world) for r in range(world)]
    max_rows = max(e - b for b, e in spans)
    padded = torch.zeros((max_rows, width_bytes), dtype=torch.uint8, device=band.device)
    padded[: band.shape[0]] = band
    out = torch.empty((world, max_rows, width_bytes), dtype=torch.uint8, device=band.device)
    dist.all_gather_into_tensor(out.view(-1), padded.view(-1), group=group)
    return torch.cat([out[r, : e - b] for r, (b, e) in enumerate(spans)], dim=0)


def max_over_ranks(value, device=None, group=None):
    """Max of a python float over ranks (device-timed durations are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
