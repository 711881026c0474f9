"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else v
    short = re.sub(r'\(CUtensor.*', '', r[ki])
    short = re.sub(r'\((const|int|float|long|__nv).*', '', short)
    short = re.sub(r'.*::', '', short)
    agg[short][0] += 1
    agg[short][1] += v
    tot += v
print('total us %.1f over %d launches' % (tot, sum(n for n, _ in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f'{t:10.1f} us {100*t/tot:5.1f}%  n={n:4d}  {k[:110]}')
