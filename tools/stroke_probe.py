#!/usr/bin/env python
"""The device stroker on a batch of large strokes: the 4096 random 64-curve paths of config 4, stroked (pen width 3, Round
joins) into 512x512 Matte8 rasters by ftl_batch_stroke.  For `ncu --metrics gpu__time_duration.sum` (kernel list of one
call) and for wall clock per call, device stroker against the host stroker (FTL_DEVICE_STROKE=0)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from footile_b200 import Batch, Format, JoinStyle, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ops, offs, _ = scenes.random_curve_paths(0, n)
pw = np.zeros(1, dtype=ops.dtype)
pw["tag"] = 5
pw["v"][0, 0] = 3.0
sops = np.insert(ops, np.asarray(offs[:-1], dtype=np.int64), pw)
soffs = np.asarray(offs, dtype=np.uint64) + np.arange(n + 1, dtype=np.uint64)
b = Batch(512, 512, Format.Matte8, n).set_join(JoinStyle.Round)
for mode in ("1", "0"):
    os.environ["FTL_DEVICE_STROKE"] = mode
    for _ in range(2):
        b.stroke(sops, soffs)
    b.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        b.stroke(sops, soffs)
    b.sync()
    dt = (time.perf_counter() - t0) / reps
    print("%s stroker: %.2f ms per batch of %d strokes (%.0f strokes/s), checksum %016x" % (
        "device" if mode == "1" else "host  ", 1e3 * dt, n, n / dt, int(np.bitwise_xor.reduce(b.checksums(0, n)))))
