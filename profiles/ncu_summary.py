#!/usr/bin/env python
"""Print the key metrics of an .ncu-rep (run in the build container: `ncu -i` needs no GPU)."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block', 'launch__waves',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct',
        'smsp__issue_active.avg.pct', 'sm__inst_executed_pipe_fp64', 'sm__pipe_fp64_cycles_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu', 'sm__pipe_tensor',
        'smsp__average_warp_latency_issue_stalled', 'smsp__average_warps_issue_stalled',
        'lts__t_sector_hit_rate', 'l1tex__t_sector_hit_rate', 'lts__t_bytes.sum ',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__inst_executed_pipe_fp64',
        'l1tex__data_bank_conflicts', 'smsp__thread_inst_executed_per_inst_executed', 'smsp__warps_eligible']


def main(path, extra=()):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('== kernel:', vals[hdr.index('Kernel Name')][:90])
        for h, u, v in zip(hdr, units, vals):
            if any(w in h for w in list(WANT) + list(extra)):
                print('  %-95s %-14s %s' % (h, u, v))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
