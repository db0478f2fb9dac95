#!/bin/bash
# K1w32 A/B: register-pinned lane constants / base pointers (MWD_W32_PIN bits: 1 = lane i/j, 2 = checkpoint base, 4 = table base)
set -u
mkdir -p gpurun_out
for tag in ${VARIANTS:-pin0 pin1 pin6 pin7}; do
  MWD_B200_LIB=$PWD/tools/scratch/libmwd_$tag.so python bench.py --no-cpu-baseline --steps 4 > gpurun_out/j_$tag.json 2> gpurun_out/j_$tag.err
  python - $tag <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/j_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'], 2), d['parity_vs_float64']['max'])
PY
done
