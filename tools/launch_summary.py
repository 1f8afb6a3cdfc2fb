#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, last and mean duration (us)."""
import collections
import csv
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split('(')[0][:64]
        v = float(r[-1].replace(',', ''))
        unit = r[-2]
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else (v * 1e6 if unit == 's' else v))
        agg.setdefault(name, []).append(v)
    print("==", path, len(rows), "launches")
    for k, v in agg.items():
        print("  %-66s n=%3d  last=%10.1f us  mean=%10.1f us" % (k, len(v), v[-1], sum(v) / len(v)))
