for cfg in "3 1 2" "3 0 2" "3 0 3" "3 0 4" "3 0 5"; do
  set -- $cfg
  echo "CTAS=$1 TAB=$2 B=$3"
  MWD_ESTEP_CTAS=$1 MWD_ESTEP_TAB=$2 MWD_ESTEP_B=$3 python bench.py --pairs 300000 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('  ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
done
