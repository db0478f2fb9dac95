#!/usr/bin/env python
"""Headline metrics + stall reasons of every kernel in an `ncu --set full` report, and (with --traffic) the
per-kernel DRAM traffic JSON that bench.py's roofline.traffic reads.

  python profiles/ncu_digest.py report.ncu-rep [--traffic out.json --pairs N --variant coco5 --concepts 65]
"""
import argparse
import csv
import io
import json
import re
import subprocess

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum']
# bench.py timer name of each kernel
TIMER = [('ik_concept', 'ik_concept'), ('ik_estep', 'ik_estep'), ('ik_counts', 'ik_estep'),
         ('posterior_grad', 'posterior_grad'), ('posterior', 'posterior')]


def to_bytes(v, unit):
    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[unit]


def to_ms(v, unit):
    return float(v) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, {'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}.get(unit, 1.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--traffic')
    ap.add_argument('--pairs', type=int, default=1000000)
    ap.add_argument('--variant', default='coco5')
    ap.add_argument('--concepts', type=int, default=65)
    ap.add_argument('--how', default='')
    a = ap.parse_args()
    out = subprocess.run(['ncu', '-i', a.report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    kernels = {}
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print('=' * 100)
        for i in idx:
            print('%-78s %-16s %s' % (hdr[i], units[i], r[i][:100]))
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and 'per_issue_active' in h and '_not_issued' not in h:
                try:
                    if float(r[i]) > 0.15:
                        print('   stall %-50s %.2f per issue' % (
                            h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(r[i])))
                except ValueError:
                    pass
        short = re.sub(r'.*::', '', name.split('(')[0]).strip()
        timer = next((t for pat, t in TIMER if pat in short), short)
        rd = to_bytes(r[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_read.sum')])
        wr = to_bytes(r[hdr.index('dram__bytes_write.sum')], units[hdr.index('dram__bytes_write.sum')])
        ms = to_ms(r[hdr.index('gpu__time_duration.sum')], units[hdr.index('gpu__time_duration.sum')])
        k = kernels.setdefault(timer, {'dram_bytes_per_launch': 0.0, 'components': {}})
        k['components'][short] = {'dram_bytes_read': rd, 'dram_bytes_write': wr, 'gpu_time_ms': ms}
        k['dram_bytes_per_launch'] += rd + wr
    if a.traffic:
        for k in kernels.values():
            k['dram_bytes_per_pair'] = k['dram_bytes_per_launch'] / a.pairs
        json.dump({'pairs_per_launch': a.pairs, 'variant': a.variant, 'n_concepts': a.concepts, 'kernels': kernels,
                   'how': a.how}, open(a.traffic, 'w'), indent=1)


if __name__ == '__main__':
    main()
