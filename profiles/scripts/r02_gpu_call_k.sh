#!/bin/bash
# K1w32 A/B: checkpoint prefetch of the backward sweep (MWD_W32_PF: 0 off, 1 prefetch.global.L1 + ld.ca, 2 registers)
set -u
mkdir -p gpurun_out
for tag in pf0 pf1 pf2; do
  export MWD_B200_LIB=$PWD/tools/scratch/libmwd_$tag.so
  python -m pytest tests/test_gpu_mixed_precision.py -x -q -k "recursion" > gpurun_out/k_tests_$tag.log 2>&1; echo "$tag pytest exit $?"
  python bench.py --no-cpu-baseline --steps 4 > gpurun_out/k_$tag.json 2> gpurun_out/k_$tag.err
  python - $tag <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/k_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'], 2), d['parity_vs_float64']['max'])
PY
done
