#!/bin/bash
# per-bucket (per n) times of the float32 vs float64 concept-chain kernels: Flickr shape (K = 100) and coco10 (K = 65)
set -u
mkdir -p gpurun_out
for v in flickr coco10; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ik_concept' -c 120 --csv \
     --log-file gpurun_out/p_concept_$v.csv python bench.py --variant $v --pairs 1000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/p_$v.log 2>&1
  echo "$v exit $?"
done
