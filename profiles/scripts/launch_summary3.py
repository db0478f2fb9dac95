#!/usr/bin/env python
"""Per-kernel totals of an ncu --csv launch list captured with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum: launches, total time, share of the
run, DRAM bytes read / written per launch.

  python profiles/scripts/launch_summary3.py launches.csv [top]
"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h = rows[hdr]
kn, mn, mu, mv = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Unit'), h.index('Metric Value')
scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}
tot = defaultdict(lambda: defaultdict(float))
cnt = defaultdict(int)
for r in rows[hdr + 1:]:
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', r[kn].split('(')[0])[:74]
    try:
        v = float(r[mv].replace(',', '')) * scale.get(r[mu], 1.0)
    except ValueError:
        continue
    tot[name][r[mn]] += v
    if r[mn] == 'gpu__time_duration.sum':
        cnt[name] += 1
T = sum(v['gpu__time_duration.sum'] for v in tot.values())
print('%-76s %4s %10s %6s %10s %11s' % ('kernel', 'n', 'ms total', 'share', 'GB read/l', 'GB write/l'))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum'])[:top]:
    n = max(cnt[k], 1)
    print('%-76s %4d %10.3f %5.1f%% %10.3f %11.3f' % (k, n, v['gpu__time_duration.sum'], 100 * v['gpu__time_duration.sum'] / T,
                                                    v['dram__bytes_read.sum'] / n, v['dram__bytes_write.sum'] / n))
