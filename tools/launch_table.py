#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/launch_table.py gpurun_out/launches.csv [...]
"""
import csv
import sys
from collections import OrderedDict

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][:64]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1].replace(",", ""))
    print(path)
    total = sum(t for _, t in agg.values())
    for k, (n, t) in agg.items():
        print("  %-66s n=%4d avg=%10.1f %s  share=%5.1f%%" % (k, n, t / n, rows[0][-2], 100.0 * t / total))
