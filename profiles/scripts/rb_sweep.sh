#!/bin/bash
# Experiment: register-block (B = 2) variant of the warp E-step vs the shared-memory block variant.
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['avg_log_likelihood'])
    elif 'rror' in l: print(l.strip())
"; }
run MWD_ESTEPW_RB=0
run MWD_ESTEPW_RB=1
run MWD_ESTEPW_RB=1 MWD_ESTEPW_OBS=0
