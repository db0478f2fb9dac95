#!/bin/bash
# float32 concept chains: pairs-per-CTA sweep (tail-warp fill) after the DMMA prologue / cheap epilogue rewrite
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mixed_precision.py -x -q -k "concept" > gpurun_out/d_tests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/d_tests.log
for ppc in 1 2 3 6; do
  MWD_CONCEPT32_PPC=$ppc python bench.py --mixed all --no-cpu-baseline --steps 3 > gpurun_out/d_bench_ppc$ppc.json 2> gpurun_out/d_bench_ppc$ppc.err
  python - $ppc <<'PY'
import json, sys
f = 'gpurun_out/d_bench_ppc%s.json' % sys.argv[1]
try:
    d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    print('ppc', sys.argv[1], d['ms_per_step'], d['kernel_ms_per_step']['ik_concept'], d['parity_vs_float64']['max'])
except Exception as e:
    print(f, 'unreadable', e)
PY
done
MWD_CONCEPT32_PPC=3 python -m pytest tests/test_gpu_mixed_precision.py -x -q -k "concept" > gpurun_out/d_tests3.log 2>&1; echo "pytest ppc3 exit $?"; tail -3 gpurun_out/d_tests3.log
