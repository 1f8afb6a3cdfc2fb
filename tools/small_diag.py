#!/usr/bin/env python
"""Diagnostic: one path through the one-launch kernel and through the general pipeline; prints edges, fill info, rasters."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from footile_b200 import FillRule, Format, Path2D, Plotter, Raster

path = (Path2D().absolute().move_to(8.0, 4.0).line_to(8.0, 3.0).cubic_to(8.0, 3.0, 8.0, 3.0, 9.0, 3.75)
        .line_to(8.0, 3.75).line_to(8.5, 3.75).line_to(8.5, 3.5).finish())
res = {}
for mode in ("small", "general"):
    if mode == "general":
        os.environ["FTL_NO_SMALL"] = "1"
    else:
        os.environ.pop("FTL_NO_SMALL", None)
    g = Plotter(Raster(16, 16, Format.Matte8))
    g.fill(FillRule.NonZero, path, (255,))
    img = g.raster().pixels
    e = g.debug_edges()
    e = e[np.lexsort(e.T[::-1])]
    print(mode, g.debug_last_fill())
    print(e)
    print(img[2:5, 6:11])
    res[mode] = img
o = oracle.Plotter(16, 16, oracle.MATTE8)
o.fill(0, path, (255,))
print("oracle", o.last_info())
print(o.raster()[2:5, 6:11])
