#!/bin/bash
# count post-pass: compact rows through ld.global.nc (L1-allocating) vs streaming loads
set -u
mkdir -p gpurun_out
for tag in base ldg base ldg; do
  if [ $tag = base ]; then unset MWD_B200_LIB; else export MWD_B200_LIB=$PWD/tools/scratch/libmwd_$tag.so; fi
  python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/z_$tag.json 2> gpurun_out/z_$tag.err
  python - $tag <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/z_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, d['parity_vs_float64']['max'])
PY
done
