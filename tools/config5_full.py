#!/usr/bin/env python
"""Config 5 at full size: one 32768x32768 Matte8 raster, 160 000 closed 64-gon sub-figures (~10.2 M edges)
in ONE fill.  Times the GPU fill, checks (a) seeded row stripes against the CPU oracle (order-free form,
u32 vertex ids), (b) that the same fill split into row bands reproduces the unsplit raster.
Usage: python tools/config5_full.py [--polys 160000] [--size 32768] [--stripes 6] [--bands 4]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from footile_b200 import FillRule, Format, Plotter, Raster, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--polys", type=int, default=160000)
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--stripes", type=int, default=6)
    ap.add_argument("--stripe-rows", type=int, default=16)
    ap.add_argument("--bands", type=int, default=4)
    ap.add_argument("--rule", type=int, default=1)
    args = ap.parse_args()
    size = args.size
    t0 = time.time()
    ops = scenes.random_polygons(0, args.polys, vertices=64, size=size, extent=2048)
    t_gen = time.time() - t0
    g = Plotter.with_clear(size, size, Format.Matte8)
    g.fill(args.rule, ops, (255,)).sync()  # warm-up: sizes the scratch buffers
    times = []
    for _ in range(3):
        t0 = time.time()
        g.fill(args.rule, ops, (255,)).sync()
        times.append(time.time() - t0)
    info = g.debug_last_fill()
    t0 = time.time()
    full = g.raster().pixels
    t_read = time.time() - t0
    px = size * (size - max(info["top_row"], 0))
    out = {"config": "one %dx%d Matte8 raster, %d sub-figures, %d edges, rule %d" % (size, size, args.polys, args.polys * 64, args.rule),
           "gen_s": t_gen, "fill_s_incl_h2d": min(times), "gpx_per_s": px / min(times) / 1e9, "edges_per_s": args.polys * 64 / min(times),
           "read_back_s": t_read, "info": info, "nonzero_pixels": int(np.count_nonzero(full))}
    # (a) stripes against the oracle
    rng = np.random.default_rng(5)
    bad = 0
    o = oracle.Plotter(size, size, oracle.MATTE8, vid_cap=1 << 30, orderfree=True)
    t0 = time.time()
    for _ in range(args.stripes):
        r0 = int(rng.integers(0, size - args.stripe_rows))
        o.set_rows(r0, r0 + args.stripe_rows)
        o.fill(args.rule, ops, (255,))
        exp = o.raster()[r0: r0 + args.stripe_rows]
        if o.last_info() != info or not np.array_equal(exp, full[r0: r0 + args.stripe_rows]):
            bad += 1
    out["oracle_stripes"] = {"n": args.stripes, "rows_each": args.stripe_rows, "mismatching": bad, "oracle_s": time.time() - t0}
    del o
    # (b) row bands reproduce the unsplit raster
    band_bad = 0
    band_times = []
    for k in range(args.bands):
        r0, r1 = k * size // args.bands, (k + 1) * size // args.bands
        gb = Plotter.with_clear(size, size, Format.Matte8, rows=(r0, r1))
        gb.fill(args.rule, ops, (255,)).sync()
        t0 = time.time()
        gb.fill(args.rule, ops, (255,)).sync()
        band_times.append(time.time() - t0)
        if not np.array_equal(gb.raster().pixels, full[r0:r1]):
            band_bad += 1
        del gb
    out["bands"] = {"n": args.bands, "mismatching": band_bad, "fill_s_each": band_times}
    print(json.dumps(out))
    return 0 if bad == 0 and band_bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
