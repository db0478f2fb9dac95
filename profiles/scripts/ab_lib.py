#!/usr/bin/env python
"""Run another script of this directory against an alternative build of libmwd_b200.so (A/B timing of a
kernel change on the SAME box, since box-to-box variance exceeds most kernel deltas):

    python profiles/scripts/ab_lib.py path/to/libmwd_b200_old.so bench_hmm.py [args...]
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import multimodalworddiscovery_b200._lib as L   # noqa: E402

L.LIB_PATH = os.path.abspath(sys.argv[1])
script = os.path.join(os.path.dirname(os.path.abspath(__file__)), sys.argv[2])
sys.argv = [script] + sys.argv[3:]
runpy.run_path(script, run_name='__main__')
