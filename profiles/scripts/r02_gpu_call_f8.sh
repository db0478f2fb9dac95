#!/bin/bash
# 8-GPU check of the NUMA-local pinned staging (round-1 e2e scaling stopped at 2.77x): topology, then the bench
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/f8_topo.txt 2>&1
lscpu | grep -i "numa\|^CPU(s)\|model name\|socket" >> gpurun_out/f8_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/f8_bench.json 2> gpurun_out/f8_bench.err
echo "bench exit $?"
tail -c 400 gpurun_out/f8_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/f8_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['e2e'])
PY
