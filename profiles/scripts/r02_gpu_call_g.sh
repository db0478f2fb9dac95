set -u
mkdir -p gpurun_out
python profiles/scripts/h2d_probe.py > gpurun_out/g_probe1.json 2> gpurun_out/g_probe1.err; echo "probe exit $?"; cat gpurun_out/g_probe1.json; tail -5 gpurun_out/g_probe1.err
python bench.py --variant coco10 --no-cpu-baseline --steps 3 > gpurun_out/g_coco10.json 2> gpurun_out/g_coco10.err; echo "coco10 exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/g_coco10.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['kernel_ms_per_step'], d['parity_vs_float64']['max'], d['float64_path']['ms_per_step'])
PY
