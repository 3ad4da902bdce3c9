#!/usr/bin/env python
"""Per-kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import defaultdict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
t, n = defaultdict(float), defaultdict(int)
for r in rows:
    k = r['Kernel Name'].split('(')[0]
    t[k] += float(r['Metric Value']); n[k] += 1
tot = sum(t.values())
print("| kernel | launches | avg us | share |\n|---|---|---|---|")
for k in sorted(t, key=lambda k: -t[k]):
    print("| %s | %d | %.1f | %.1f%% |" % (k, n[k], t[k] / n[k] / 1e3, 100 * t[k] / tot))
