#!/usr/bin/env python
"""Where a single small fill spends its time: ftl_fill + ftl_sync of benches/fishyb.rs fill_256 in a loop, for
`ncu --metrics gpu__time_duration.sum` (kernel list of one call) and for wall-clock per call with and without the graph."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from footile_b200 import FillRule, Format, Plotter, Raster, scenes

path = scenes.fishy_bench()
g = Plotter(Raster(256, 256, Format.Matte8)).set_transform([2, 0, 0, 0, 2, 0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for _ in range(5):
    g.fill(FillRule.NonZero, path, (255,)).sync()
t0 = time.perf_counter()
for _ in range(n):
    g.fill(FillRule.NonZero, path, (255,)).sync()
t1 = time.perf_counter()
for _ in range(n):
    g.fill(FillRule.NonZero, path, (255,))
g.sync()
t2 = time.perf_counter()
print("fill+sync %.1f us per call; %d fills then one sync %.1f us per call" % ((t1 - t0) / n * 1e6, n, (t2 - t1) / n * 1e6))
