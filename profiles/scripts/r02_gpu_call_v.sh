#!/bin/bash
# compact (float32 mantissa + exponent) row statistics between the float32 recursion kernel and the count post-pass:
# parity (goldens, mixed gate, float64 suites that share the count kernels), A/B, DRAM bytes of the two kernels
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/v_tests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/v_tests.log
for tag in wide new; do
  if [ $tag = new ]; then unset MWD_B200_LIB; else export MWD_B200_LIB=$PWD/tools/scratch/libmwd_$tag.so; fi
  python bench.py --no-cpu-baseline --steps 4 > gpurun_out/v_$tag.json 2> gpurun_out/v_$tag.err
  python - $tag <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/v_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
pv = d['parity_vs_float64']
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, {k: v for k, v in pv.items() if k != 'what'})
PY
done
unset MWD_B200_LIB
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k regex:'ik_estep_warp32|ik_counts_small' -c 2 --csv --log-file gpurun_out/v_ncu_new.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep "ik_estep_warp32\|ik_counts" gpurun_out/v_ncu_new.csv | awk -F'","' '{print "   ", $5, $(NF-2), $(NF-1), $NF}' | cut -c1-200
