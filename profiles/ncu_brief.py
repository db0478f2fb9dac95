#!/usr/bin/env python
"""Brief digest of an ncu report for one kernel: headline metrics + per-opcode executed
instructions / shared wavefronts per unit of work.

  python profiles/ncu_brief.py report.ncu-rep <kernel-regex> [units]   (units: e.g. pair-steps)
"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter


def ncu_csv(rep, page, kern):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv', '--kernel-name', 'regex:' + kern],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, kern, units=1.0):
    rows = ncu_csv(rep, 'raw', kern)
    hdr, vals = rows[0], rows[2]
    want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
            'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
            'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
            'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
    for h, u, v in zip(hdr, rows[1], vals):
        if h in want or 'issue_stalled' in h and 'per_issue_active' in h:
            print('%-86s %-10s %s' % (h, u, v))
    rows = ncu_csv(rep, 'source', kern)
    hdr = rows[1]
    ci, src, ws = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('L1 Wavefronts Shared')
    tg = hdr.index('L1 Tag Requests Global')
    c, w, g = Counter(), Counter(), Counter()
    for r in rows[2:]:
        if len(r) <= ci:
            continue
        try:
            n = int(r[ci] or 0)
        except ValueError:
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[src].strip())
        op = '.'.join((m.group(2) if m else r[src][:10]).split('.')[:3])
        c[op] += n
        w[op] += int(r[ws] or 0)
        g[op] += int(r[tg] or 0)
    tot = sum(c.values())
    print('warp-instructions per unit: %.1f' % (tot / units))
    for op, n in c.most_common(28):
        print('  %-22s %5.1f%%  %8.1f /unit   shared-wavefronts %7.1f   global-tags %7.1f'
              % (op, 100.0 * n / tot, n / units, w[op] / units, g[op] / units))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
