#!/bin/bash
# Experiment: concept chains on a side stream next to the recursion kernel (launched first) with
# 3/2/1 recursion CTAs per SM.
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['avg_log_likelihood'])
"; }
run MWD_OVERLAP=1 MWD_ESTEPW_CTAS=3
run MWD_OVERLAP=1 MWD_ESTEPW_CTAS=2
run MWD_OVERLAP=0 MWD_ESTEPW_CTAS=2
run MWD_OVERLAP=1 MWD_ESTEPW_CTAS=1
