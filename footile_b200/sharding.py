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
    """Contiguous block partition of range(n): returns (first, count) of this rank."""
    base, extra = divmod(int(n), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def band_rows(height, rank, world, align=8):
    """Row band [begin, end) of rank `rank`; band boundaries are multiples of `align` rows."""
    units = (int(height) + align - 1) // align
    first, count = shard_range(units, rank, world)
    return min(first * align, height), min((first + count) * align, height)


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
