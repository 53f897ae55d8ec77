"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys

path = sys.argv[1]
lines = open(path).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
agg = collections.OrderedDict()
for r in rows:
    k = r['Kernel Name'][:70]
    a = agg.setdefault(k, [0, 0.0, r['Block Size'], r['Grid Size']])
    a[0] += 1
    a[1] += float(r['Metric Value'])
tot = sum(a[1] for a in agg.values())
print('%d launches, %.1f us total' % (len(rows), tot / 1e3))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print('%-72s n=%4d total %10.1f us  avg %9.1f us  %5.1f%%  block %s grid %s' % (
        k, a[0], a[1] / 1e3, a[1] / 1e3 / a[0], 100 * a[1] / tot, a[2], a[3]))
