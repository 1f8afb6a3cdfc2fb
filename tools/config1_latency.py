#!/usr/bin/env python
"""Config 1 (benches/fishyb.rs:10-39): fill_16 / fill_256 / stroke_16 / stroke_256 on Matte8 with scale(2,2),
plus the examples/fishy.rs scene on 128x128 Rgba8p.  Per-call latency of the GPU path through the C ABI
(call + ftl_sync, raster resident on the device) beside the CPU oracle (one thread), both including the
raster allocation only where the reference bench includes it (it does: fishyb.rs:18-20,34-39) -> reported
both ways.  This is a latency config, not a bandwidth one (<= 64 KiB of pixels)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from footile_b200 import FillRule, Format, Plotter, Raster, scenes  # noqa: E402

T = [2, 0, 0, 0, 2, 0]


def timeit(fn, n):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    path = scenes.fishy_bench()
    out = {}
    for size in (16, 256):
        g = Plotter(Raster(size, size, Format.Matte8)).set_transform(T)
        o = oracle.Plotter(size, size, oracle.MATTE8).set_transform(T)
        out["fill_%d" % size] = {
            "gpu_us_per_call_resident": timeit(lambda: g.fill(FillRule.NonZero, path, (255,)).sync(), 300),
            "gpu_us_incl_new_plotter_and_readback": timeit(lambda: Plotter(Raster(size, size, Format.Matte8)).set_transform(T).fill(0, path, (255,)).raster(), 50),
            "cpu_us_per_call": timeit(lambda: o.fill(0, path, (255,)), 300),
            "cpu_us_incl_new_plotter": timeit(lambda: oracle.Plotter(size, size, oracle.MATTE8).set_transform(T).fill(0, path, (255,)), 300)}
        out["stroke_%d" % size] = {
            "gpu_us_per_call_resident": timeit(lambda: g.stroke(path, (255,)).sync(), 200),
            "cpu_us_per_call": timeit(lambda: o.stroke(path, (255,)), 300)}
    fish, eye = scenes.fishy_example()
    g = Plotter(Raster(128, 128, Format.Rgba8p))
    o = oracle.Plotter(128, 128, oracle.RGBA8P)

    def scene(p):
        p.fill(0, fish, (127, 96, 96, 255))
        p.stroke(fish, (255, 208, 208, 255))
        p.stroke(eye, (0, 0, 0, 255))

    out["fishy_example_128_rgba8p"] = {"gpu_us": timeit(lambda: (scene(g), g.sync()), 100), "cpu_us": timeit(lambda: scene(o), 100)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
