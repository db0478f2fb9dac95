#!/bin/bash
# 8 GPUs: which staging memory feeds all GPUs fastest (profiles/scripts/h2d_probe.py)
set -u
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled > gpurun_out/h8_thp.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
    profiles/scripts/h2d_probe.py > gpurun_out/h8_probe.json 2> gpurun_out/h8_probe.err
echo "probe exit $?"; grep '^{' gpurun_out/h8_probe.json; tail -3 gpurun_out/h8_probe.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 \
    profiles/scripts/h2d_probe.py > gpurun_out/h8_probe4.json 2> gpurun_out/h8_probe4.err
echo "probe4 exit $?"; grep '^{' gpurun_out/h8_probe4.json
