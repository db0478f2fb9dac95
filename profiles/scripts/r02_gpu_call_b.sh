#!/bin/bash
# concept chains in float32 with the (hi, lo) emission pair and paired chains per lane: tests, error trajectory, timing
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mixed_precision.py -x -q -k "concept" > gpurun_out/b_tests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/b_tests.log
python profiles/scripts/mixed_trajectory.py concept all > gpurun_out/b_traj.txt 2>&1; echo "traj exit $?"; tail -4 gpurun_out/b_traj.txt
python bench.py --mixed all --no-cpu-baseline --steps 3 > gpurun_out/b_bench_all.json 2> gpurun_out/b_bench_all.err; echo "bench exit $?"
python bench.py --mixed all --variant coco10 --no-cpu-baseline --steps 3 > gpurun_out/b_bench_all_coco10.json 2> gpurun_out/b_bench_all_coco10.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/b_bench_all.json', 'gpurun_out/b_bench_all_coco10.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['ms_per_step'], d['kernel_ms_per_step'], d['parity_vs_float64'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
