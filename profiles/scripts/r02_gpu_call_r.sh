#!/bin/bash
# K1w32 at 5 CTAs per SM (96 registers, 20 partial rows per SM) vs the default 4; align leg per pass
set -u
mkdir -p gpurun_out
for tag in base 5cta; do
  if [ $tag = base ]; then unset MWD_B200_LIB; else export MWD_B200_LIB=$PWD/tools/scratch/libmwd_$tag.so; fi
  python bench.py --no-cpu-baseline --steps 4 > gpurun_out/r_$tag.json 2> gpurun_out/r_$tag.err
  python - $tag <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/r_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'], 2), d['parity_vs_float64']['max'])
print('   align', [round(x, 2) for x in d['align']['ms_each_pass']])
PY
done
