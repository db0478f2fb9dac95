#!/bin/bash
# after the padded emission rows in both warp kernels: full GPU suite, float64 and mixed bench lines
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/m_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/m_gputest.log
tail -3 gpurun_out/m_gputest.log
python bench.py --no-cpu-baseline --steps 4 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/m_bench.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'], 2), d['parity_vs_float64']['max'])
print('float64', d['float64_path']['ms_per_step'], d['float64_path']['kernel_ms_per_step'])
PY
