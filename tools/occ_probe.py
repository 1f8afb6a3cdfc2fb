#!/usr/bin/env python
"""Tile-kernel time of a store-only (analytic) Matte8 launch against the number of resident CTAs per SM,
for several raster sizes: the evidence behind the 3-CTA policy in Engine::replay()."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import footile_b200 as fb
from footile_b200 import Batch, Format, Path2D


def star(size):
    c, r = size / 2.0, 0.45 * size
    p = Path2D().absolute()
    for n in range(7):
        t = np.float32(4.0 * math.pi * n / 7.0)
        x, y = float(c + r * np.cos(t)), float(c + r * np.sin(t))
        p = p.move_to(x, y) if n == 0 else p.line_to(x, y)
    return p.close().finish()


for size in (1024, 2048, 4096, 8192):
    n = max(4, (1 << 30) // (size * size))
    ops, offs = Batch.pack([star(size)] * n)
    b = Batch(size, size, Format.Matte8, n)
    b.upload(ops, offs)
    res = []
    for occ in (2, 3, 4, 5):
        os.environ["FTL_OCC"] = str(occ)
        for _ in range(3):
            b.run()
        b.sync()
        fb.set_profiling(True)
        fb.tile_kernel_time(reset=True)
        for _ in range(10):
            b.run()
        b.sync()
        ms, k = fb.tile_kernel_time(reset=True)
        fb.set_profiling(False)
        res.append("occ %d: %.3f ms" % (occ, ms / k))
    print("%5d^2 x %4d: %s" % (size, n, "  ".join(res)))
