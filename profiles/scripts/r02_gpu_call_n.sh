#!/bin/bash
# generic-width warp kernels: full GPU suite, then K = 57 (no exact instantiation) at 1 M pairs with the warp kernels
# vs the CTA-per-4-pairs fallback, float64 and mixed
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/n_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/n_gputest.log
tail -3 gpurun_out/n_gputest.log
export MWD_BENCH_CONCEPTS=57
for tag in gen cta; do
  if [ $tag = cta ]; then export MWD_ESTEP_WARP=0 MWD_ESTEP_WARP32=0; fi
  python bench.py --no-cpu-baseline --steps 3 > gpurun_out/n_k57_$tag.json 2> gpurun_out/n_k57_$tag.err
  python - $tag <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/n_k57_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, d['parity_vs_float64']['max'])
print('   float64', round(d['float64_path']['ms_per_step'], 3), {k: round(v, 3) for k, v in d['float64_path']['kernel_ms_per_step'].items()})
PY
done
