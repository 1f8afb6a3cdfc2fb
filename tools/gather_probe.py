#!/usr/bin/env python
"""The optional final gather of row bands (no collective on the data path; only this): every rank holds its band
of a 32768 x 32768 Matte8 raster (1 GiB in all) and all-gathers the full raster over NCCL / NVLink.  Run under torchrun."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from footile_b200 import sharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
size = 32768
r0, r1 = sharding.band_rows(size, rank, world, align=32)
band = torch.full((r1 - r0, size), rank + 1, dtype=torch.uint8, device="cuda")
for _ in range(2):
    full = sharding.gather_bands(band, size, size)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 5
for _ in range(reps):
    full = sharding.gather_bands(band, size, size)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ok = all(int(full[sharding.band_rows(size, r, world, align=32)[0], 0]) == r + 1 for r in range(world)) and full.shape == (size, size)
if rank == 0:
    print("gather of %d bands of a %dx%d Matte8 raster (all_gather + concatenation on every rank): %.2f ms per gather (max over ranks), "
          "%.0f GB/s received per rank, contents ok: %s" % (world, size, size, float(ms), size * size * (world - 1) / world / (float(ms) * 1e-3) / 1e9, ok))
dist.destroy_process_group()
