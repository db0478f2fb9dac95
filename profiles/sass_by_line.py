#!/usr/bin/env python
"""Join an ncu SASS-page CSV (per-instruction 'Instructions Executed' / stall samples) with
nvdisasm -g line info of the same kernel and aggregate per CUDA source line.

  ncu -i rep.ncu-rep --page source --csv > sass.csv
  cuobjdump -xelf all lib.so ; nvdisasm -g -c file.cubin > lines.txt
  python profiles/sass_by_line.py sass.csv lines.txt <mangled-kernel-substring> [source.cu]
"""
import csv
import re
import sys
from collections import defaultdict


def main(sass_csv, lines_txt, kernel, src=None, top=45):
    rows = list(csv.reader(open(sass_csv)))
    # keep the first kernel section only (a report with several launches repeats the header)
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    if len(starts) > 1:
        rows = rows[:starts[1]]
    hdr = rows[1]
    ci = hdr.index('Instructions Executed')
    cs = hdr.index('# Samples')
    insts = [(r[1].strip(), int(r[ci] or 0), int(r[cs] or 0)) for r in rows[2:] if len(r) > ci]
    # line info in instruction order for the kernel's section
    cur = None
    lines = []
    active = False
    for ln in open(lines_txt):
        if ln.startswith('//---') and '.text.' in ln:
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
            lines.append(cur)
    if len(lines) != len(insts):
        print('warning: %d disassembled vs %d profiled instructions' % (len(lines), len(insts)))
    per = defaultdict(lambda: [0, 0, 0])
    for (txt, n, smp), line in zip(insts, lines):
        per[line][0] += n
        per[line][1] += smp
        per[line][2] += 1
    tot = sum(v[0] for v in per.values()) or 1
    tots = sum(v[1] for v in per.values()) or 1
    import os
    srcl = open(src).read().splitlines() if src else None
    srcname = os.path.basename(src) if src else None
    print('total warp-instructions %d, samples %d, static SASS %d' % (tot, tots, len(insts)))
    for line, (n, smp, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ''
        if line and srcl and line[0] == srcname and line[1] <= len(srcl):
            text = srcl[line[1] - 1].strip()[:90]
        line = '%s:%d' % (line[0][:14], line[1]) if line else '?'
        print('%20s  inst %5.1f%%  stall-samples %5.1f%%  sass %4d | %s' % (line, 100.0 * n / tot, 100.0 * smp / tots, st, text))


if __name__ == '__main__':
    main(*sys.argv[1:5])
