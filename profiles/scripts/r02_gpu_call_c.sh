#!/bin/bash
# FMA-form issue rates on this B200, then one full ncu capture of the float32 concept-chain kernel
set -u
mkdir -p gpurun_out
true
CMD="python bench.py --mixed all --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/c_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ik_concept32_kernel' -c 1 \
    -f -o gpurun_out/r02_concept32 $CMD > gpurun_out/c_ncu.log 2>&1
echo "ncu exit $?"
