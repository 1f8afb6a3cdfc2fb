#!/usr/bin/env python
"""Longer randomized parity run than the test suite: the analytic-row scenes, random polygons (through the small fill, the
bins and raster_tiles' own scatter), curved paths, edge records, layered scenes and the device stroker's outlines of
tests/test_gpu_parity.py with fresh seeds.  Usage: python tools/fuzz_parity.py [seed0] [rounds]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import test_gpu_parity as T
from footile_b200 import Format

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
real_rng = np.random.default_rng
n = 0
for r in range(rounds):
    # the tests seed their generators with small constants: shift them
    np.random.default_rng = lambda s=None, _r=r: real_rng((0 if s is None else int(s)) + seed0 + 7919 * _r)
    for fmt in (Format.Matte8, Format.Rgba8p, Format.Graya8p):
        T.test_analytic_rows_vs_oracle(fmt)
        n += 1
        for env in ({}, {"FTL_NO_SMALL": "1", "FTL_DIRECT_MAX": "8"}, {"FTL_NO_SMALL": "1", "FTL_DIRECT_MAX": "64"}):
            os.environ.update(env)
            try:
                T._random_polygons_vs_oracle(fmt)
            finally:
                for k in env:
                    os.environ.pop(k, None)
            n += 1
    for rule in (0, 1):
        T.test_curved_paths_vs_oracle(rule)
        n += 1
    for seed in range(4):
        T.test_edges_bit_exact(seed)
        T.test_flatten_vertices_bit_exact(seed)
        n += 2
    for fmt in (Format.Rgba8p, Format.Matte8, Format.Graya8p):
        T.test_fill_layers_equals_sequential_calls(fmt)
        n += 1
    T.test_device_stroker_outline_equals_host_on_random_paths()
    n += 1
    print("round", r, "ok", flush=True)
np.random.default_rng = real_rng
print("fuzz: %d randomized test bodies passed with seeds from %d" % (n, seed0))
