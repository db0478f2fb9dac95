#!/bin/bash
# Experiment: register-block (B = 2) variant of the warp E-step vs the shared-memory block variant,
# on the mixed-n workload (coco10) and the headline one (coco5).
run() { echo "== $*"; env "${@:2}" python bench.py --variant $1 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['avg_log_likelihood'])
    elif 'rror' in l: print(l.strip())
"; }
run coco10 MWD_ESTEPW_RB=1
run coco10 MWD_ESTEPW_RB=0
