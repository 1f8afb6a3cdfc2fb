#!/usr/bin/env python
"""A scene of many paths on ONE raster: N random curve paths composited in order onto a
3840x2160 Rgba8p raster — one ftl_fill_layers call vs N ftl_fill calls vs the CPU oracle (N fills)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from footile_b200 import Format, Plotter, Raster, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=400)
    ap.add_argument("--cpu-layers", type=int, default=20)
    args = ap.parse_args()
    w, h = 3840, 2160
    ops, offs, rules = scenes.random_curve_paths(0, args.layers, segments=24, size=2048)
    rng = np.random.default_rng(1)
    colors = rng.integers(0, 256, (args.layers, 4)).astype(np.uint8)
    colors[:, :3] = np.minimum(colors[:, :3], colors[:, 3:4])
    layers = [(int(rules[j]), ops[int(offs[j]): int(offs[j + 1])], colors[j]) for j in range(args.layers)]
    base = Raster.with_color(w, h, Format.Rgba8p, (64, 128, 64, 255))
    a = Plotter(base)
    a.fill_layers(layers).sync()
    a.write_raster(base.pixels)
    t0 = time.perf_counter()
    a.fill_layers(layers).sync()
    t_layered = time.perf_counter() - t0
    b = Plotter(base)
    for r, o, c in layers[:5]:
        b.fill(r, o, c)
    b.sync()
    b.write_raster(base.pixels)
    t0 = time.perf_counter()
    for r, o, c in layers:
        b.fill(r, o, c)
    b.sync()
    t_seq = time.perf_counter() - t0
    same = bool(np.array_equal(a.raster().pixels, b.raster().pixels))
    o = oracle.Plotter(w, h, oracle.RGBA8P, init=base.pixels)
    t0 = time.perf_counter()
    for r, p, c in layers[: args.cpu_layers]:
        o.fill(r, p, c)
    t_cpu = (time.perf_counter() - t0) / args.cpu_layers * args.layers
    ok_prefix = None
    if args.cpu_layers == args.layers:
        ok_prefix = bool(np.array_equal(o.raster(), a.raster().pixels))
    print(json.dumps({"layers": args.layers, "raster": "%dx%d Rgba8p" % (w, h), "gpu_layered_ms": 1e3 * t_layered, "gpu_sequential_ms": 1e3 * t_seq,
                      "layered_equals_sequential": same, "cpu_oracle_ms_extrapolated": 1e3 * t_cpu, "cpu_layers_timed": args.cpu_layers,
                      "oracle_equal": ok_prefix, "layers_per_s_layered": args.layers / t_layered}))


if __name__ == "__main__":
    main()
