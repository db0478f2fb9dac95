#!/bin/bash
# Round-2 closing records of the final build: c5 / c1 / c2 / c3 bench lines and the ncu launch list of the default command
set -u
mkdir -p gpurun_out/final2
O=gpurun_out/final2
python bench.py > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 exit $?"
for c in c1 c2 c3; do
  python bench.py --config $c --steps 10 > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c exit $?"
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 \
    --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list exit $?"
python - <<'PY'
import json
for c in ('c5', 'c1', 'c2', 'c3'):
    d = json.loads([l for l in open('gpurun_out/final2/bench_%s.json' % c) if l.startswith('{')][-1])
    print(c, round(d['ms_per_step'], 3), round(d['value']), round(d['e2e']['value']), {k: round(v, 3) for k, v in d['kernel_ms_per_step'].items()}, d['parity_vs_float64']['max'], (d['cpu_baseline'] or {}).get('value'))
PY
