#!/usr/bin/env python
"""Per-call latency of small fills through the C ABI (timed inside the library, ftl_time_fills): the workloads of
benches/fishyb.rs (fill_16, fill_256 with scale(2,2) on Matte8) and the fill of examples/fishy.rs (128x128 Rgba8p),
with a sync after every call and back to back, for the one-launch path and for the general pipeline."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from footile_b200 import FillRule, Format, Plotter, Raster, scenes

out = {}
path = scenes.fishy_bench()
fish, eye = scenes.fishy_example()
for mode in ("small", "general"):
    if mode == "general":
        os.environ["FTL_NO_SMALL"] = "1"
    else:
        os.environ.pop("FTL_NO_SMALL", None)
    for size in (16, 256):
        g = Plotter(Raster(size, size, Format.Matte8)).set_transform([2, 0, 0, 0, 2, 0])
        g.time_fills(FillRule.NonZero, path, (255,), iters=50)
        out["%s_fill_%d" % (mode, size)] = {"us_sync_each": g.time_fills(FillRule.NonZero, path, (255,), iters=2000, sync_each=True),
                                            "us_back_to_back": g.time_fills(FillRule.NonZero, path, (255,), iters=2000, sync_each=False)}
    g = Plotter(Raster(128, 128, Format.Rgba8p))
    g.time_fills(FillRule.NonZero, fish, (127, 96, 96, 255), iters=50)
    out["%s_fishy_fill_128_rgba8p" % mode] = {"us_sync_each": g.time_fills(FillRule.NonZero, fish, (127, 96, 96, 255), iters=2000, sync_each=True),
                                              "us_back_to_back": g.time_fills(FillRule.NonZero, fish, (127, 96, 96, 255), iters=2000, sync_each=False)}
print(json.dumps(out))
if os.environ.get("FTL_SMALL_PROF") == "1":
    import ctypes as C
    import numpy as np
    from footile_b200 import _lib
    os.environ.pop("FTL_NO_SMALL", None)
    for size in (16, 256):
        g = Plotter(Raster(size, size, Format.Matte8)).set_transform([2, 0, 0, 0, 2, 0])
        for _ in range(5):
            g.fill(FillRule.NonZero, path, (255,)).sync()
        st = np.zeros(9, dtype=np.int64)
        _lib.check(_lib.lib().ftl_debug_small_profile(g._handle, st.ctypes.data))
        print("phases(cycles) fill_%d" % size, list(np.diff(st)), "total", int(st[8] - st[0]), file=sys.stderr)
