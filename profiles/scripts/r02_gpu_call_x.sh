#!/bin/bash
# K1w32: longest captions first vs ascending order (A/B on one box) + the recursion parity tests
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mixed_precision.py tests/test_gpu_graph.py tests/test_gpu_full_size.py -x -q > gpurun_out/x_tests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/x_tests.log
for tag in asc new asc new; do
  if [ $tag = new ]; then unset MWD_B200_LIB; else export MWD_B200_LIB=$PWD/tools/scratch/libmwd_$tag.so; fi
  python bench.py --no-cpu-baseline --steps 4 > gpurun_out/x_$tag.json 2> gpurun_out/x_$tag.err
  python - $tag <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/x_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, d['parity_vs_float64']['max'])
PY
done
