#!/bin/bash
# full ncu capture (source-level) of the float32 concept-chain kernel and the float32 recursion kernel as built now
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/l_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ik_concept32_kernel|ik_estep_warp32_kernel' -c 2 \
    -f -o gpurun_out/r02_l_full $CMD > gpurun_out/l_ncu.log 2>&1
echo "ncu exit $?"
