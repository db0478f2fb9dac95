#!/bin/bash
# float32 phone-table prologue of the float32 concept chains vs the float64 prologue: parity gate, then A/B
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mixed_precision.py tests/test_gpu_graph.py -x -q > gpurun_out/u_tests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/u_tests.log
python profiles/scripts/mixed_trajectory.py concept mixed > gpurun_out/u_traj.txt 2>&1; tail -3 gpurun_out/u_traj.txt
for tag in p64 new p64 new; do
  if [ $tag = new ]; then unset MWD_B200_LIB; else export MWD_B200_LIB=$PWD/tools/scratch/libmwd_$tag.so; fi
  python bench.py --no-cpu-baseline --steps 4 > gpurun_out/u_$tag.json 2> gpurun_out/u_$tag.err
  python - $tag <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/u_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, d['parity_vs_float64'])
PY
done
