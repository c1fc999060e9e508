#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised:
compare shares, not absolutes)."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
    tot[name][0] += 1
    tot[name][1] += us
total = sum(v[1] for v in tot.values())
n = sum(v[0] for v in tot.values())
print('# %d launches, %.1f ms summed kernel time (serialised under ncu)' % (n, total / 1e3))
for name, (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print('%10.0f us %5.1f%% n=%5d avg=%8.1f  %s' % (us, 100 * us / total, cnt, us / cnt, name[:90]))
