#!/bin/bash
# e2e leg: chunks of the streamed iteration (the kernels of the last chunk are the exposed tail)
set -u
mkdir -p gpurun_out
for ch in 32 64 128; do
  python bench.py --no-cpu-baseline --steps 4 --chunks $ch > gpurun_out/w_ch$ch.json 2> gpurun_out/w_ch$ch.err
  python - $ch <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/w_ch%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print('chunks', sys.argv[1], 'resident', round(d['ms_per_step'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'GB/s', round(d['e2e']['h2d_gbs_per_rank'], 2), 'launches', d['gpu_launches'])
PY
done
