#!/bin/bash
# A/B on one box: K1w32 with padded emission rows + uniform conceptCountsA branch vs the previous build; e2e with 16 vs 32 chunks
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mixed_precision.py -x -q -k "recursion or twenty" > gpurun_out/i_tests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/i_tests.log
for tag in base new base new; do
  if [ $tag = base ]; then export MWD_B200_LIB=$PWD/tools/scratch/libmwd_base.so; CH="--chunks 16"; else unset MWD_B200_LIB; CH=""; fi
  python bench.py --no-cpu-baseline --steps 4 $CH > gpurun_out/i_$tag.json 2> gpurun_out/i_$tag.err
  python - $tag <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/i_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'], 2), d['parity_vs_float64']['max'])
PY
done
