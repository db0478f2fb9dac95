#!/bin/bash
# per-bucket choice between the float32 and the float64 concept-chain kernels: tests, Flickr / coco10 / coco5 at 1 M pairs
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mixed_precision.py tests/test_gpu_graph.py -x -q > gpurun_out/q_tests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/q_tests.log
for spec in "flickr" "coco10" "coco5"; do
  python bench.py --variant $spec --pairs 1000000 --no-cpu-baseline --steps 3 > gpurun_out/q_$spec.json 2> gpurun_out/q_$spec.err
  python - $spec <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/q_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, d['parity_vs_float64']['max'])
print('   float64', round(d['float64_path']['ms_per_step'], 3), {k: round(v, 3) for k, v in d['float64_path']['kernel_ms_per_step'].items()})
PY
done
