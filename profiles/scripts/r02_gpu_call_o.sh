#!/bin/bash
# generic-width warp kernels, dense width list: full GPU suite, K = 57 (coco5) and K = 90 (coco10) at 1 M pairs
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/o_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/o_gputest.log
tail -3 gpurun_out/o_gputest.log
for spec in "57 coco5" "90 coco10"; do
  set -- $spec
  MWD_BENCH_CONCEPTS=$1 python bench.py --variant $2 --no-cpu-baseline --steps 3 > gpurun_out/o_k$1.json 2> gpurun_out/o_k$1.err
  python - $1 <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/o_k%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print('K', sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, d['parity_vs_float64']['max'])
print('   float64', round(d['float64_path']['ms_per_step'], 3), {k: round(v, 3) for k, v in d['float64_path']['kernel_ms_per_step'].items()})
PY
done
