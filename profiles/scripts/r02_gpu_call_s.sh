#!/bin/bash
# K1w32: L2 evict_last policy on checkpoint stores (bit 0) / phone-table updates (bit 1): time and DRAM bytes per launch
set -u
mkdir -p gpurun_out
for tag in base keep1 keep2 keep3; do
  if [ $tag = base ]; then unset MWD_B200_LIB; else export MWD_B200_LIB=$PWD/tools/scratch/libmwd_$tag.so; fi
  python bench.py --no-cpu-baseline --steps 3 > gpurun_out/s_$tag.json 2> gpurun_out/s_$tag.err
  python - $tag <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/s_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, d['parity_vs_float64']['max'])
PY
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none \
     -k regex:ik_estep_warp32 -c 1 --csv --log-file gpurun_out/s_ncu_$tag.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  grep "ik_estep_warp32" gpurun_out/s_ncu_$tag.csv | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $NF}'
done
