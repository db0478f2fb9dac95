#!/bin/bash
# 'mixed' now includes the float32 concept chains (pair emission, auto pairs-per-CTA): full GPU suite, default bench
# line, 20-iteration error trajectories (linear: done in call b; Gaussian class here)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/e_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/e_gputest.log
tail -3 gpurun_out/e_gputest.log
python bench.py > gpurun_out/e_bench_c5.json 2> gpurun_out/e_bench_c5.err; echo "bench exit $?"
tail -c 300 gpurun_out/e_bench_c5.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/e_bench_c5.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['kernel_ms_per_step'], d['parity_vs_float64']['max'], d['e2e']['value'], d['cpu_baseline'])
PY
