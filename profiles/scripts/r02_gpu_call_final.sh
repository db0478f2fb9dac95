#!/bin/bash
# Round-2 final records: every BASELINE config through bench.py, the reference (CPU) arm, the 1 M-pair variants,
# ncu launch list + one full capture of the four dominant kernels of the default bench command.
set -u
mkdir -p gpurun_out/final
O=gpurun_out/final
python bench.py > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 exit $?"
for c in c1 c2 c3 c4; do
  python bench.py --config $c --steps 10 > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c exit $?"
done
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_c5_reference_arm.json 2> $O/bench_c5_reference_arm.err; echo "reference arm exit $?"
python bench.py --variant coco10 --no-cpu-baseline --steps 3 > $O/bench_coco10.json 2> $O/bench_coco10.err; echo "coco10 exit $?"
python bench.py --variant flickr --pairs 1000000 --no-cpu-baseline --steps 3 > $O/bench_flickr1m.json 2> $O/bench_flickr1m.err; echo "flickr exit $?"
python bench.py --model gaussian --no-cpu-baseline --steps 3 > $O/bench_gaussian1m.json 2> $O/bench_gaussian1m.err; echo "gaussian exit $?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 \
    --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'ik_concept32_kernel|ik_estep_warp32_kernel|posterior_tc_kernel|posterior_grad_tc_kernel' -c 4 \
    -f -o $O/r02_final_full $CMD > $O/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la $O
