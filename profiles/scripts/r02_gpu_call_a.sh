#!/bin/bash
# Round-2 record call: GPU test suite, default bench line, ncu launch list and one full capture of the four
# dominant kernels of the same bench command (each ncu run only after the plain command exited 0).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/gputest.log
tail -3 gpurun_out/gputest.log
python bench.py > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench_c5.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 \
    --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'ik_concept_kernel|ik_estep_warp32_kernel|posterior_tc_kernel|posterior_grad_tc_kernel' -c 4 \
    -f -o gpurun_out/r02_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out
