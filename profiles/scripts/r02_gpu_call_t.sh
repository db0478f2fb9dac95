#!/bin/bash
# final sanity: smoke(), full GPU suite, default bench line
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/t_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/t_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_gputest.log
tail -3 gpurun_out/t_gputest.log
python bench.py > gpurun_out/t_bench_c5.json 2> gpurun_out/t_bench_c5.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/t_bench_c5.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, 'e2e', round(d['e2e']['value']), d['parity_vs_float64']['max'], d['roofline']['traffic'], [round(x, 2) for x in d['align']['ms_each_pass']])
PY
