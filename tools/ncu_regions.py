#!/usr/bin/env python
"""Regions of equal execution count in the SASS of a kernel (ncu source page): where the warp-instructions go."""
import csv
import subprocess
import sys

path = sys.argv[1]
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iss, ith = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
data = rows[2:]
tot = sum(int(r[ia]) for r in data)
smp = sum(int(r[iss]) for r in data)
print("total warp-instructions", tot, "SASS lines", len(data), "samples", smp)
prev, start, acc, sacc, regions = None, 0, 0, 0, []
for k, r in enumerate(data):
    c = int(r[ia])
    if prev is None or abs(c - prev) > 0.02 * max(c, prev, 1):
        if prev is not None:
            regions.append((start, k - 1, prev, acc, sacc))
        start, acc, sacc = k, 0, 0
    acc += c
    sacc += int(r[iss])
    prev = c
regions.append((start, len(data) - 1, prev, acc, sacc))
for s, e, c, a, sa in regions:
    if a > tot * minshare:
        print("%4d-%4d n=%3d count=%12d inst=%5.1f%% smp=%5.1f%% thr=%-5s %s" % (s, e, e - s + 1, c, 100 * a / tot, 100 * sa / max(smp, 1), data[s][ith], data[s][isrc].strip()[:48]))
if len(sys.argv) > 3:
    a, b = int(sys.argv[3]), int(sys.argv[4])
    for k in range(a, b + 1):
        r = data[k]
        print("%5d %12s thr=%-5s smp=%-7s %s" % (k, r[ia], r[ith], r[iss], r[isrc].strip()[:100]))
