#!/usr/bin/env python
"""Sum gpu__time_duration per kernel name of an ncu --csv launch list."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h = rows[hdr]
kn, mv = h.index('Kernel Name'), h.index('Metric Value')
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[hdr + 1:]:
    try:
        tot[r[kn].split('(')[0][:70]] += float(r[mv].replace(',', ''))
        cnt[r[kn].split('(')[0][:70]] += 1
    except ValueError:
        pass
T = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print('%-72s n=%4d %10.3f ms %5.1f%%' % (k, cnt[k], v / 1e6, 100 * v / T))
