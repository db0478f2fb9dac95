#!/usr/bin/env python
"""Stage an UNTRACKED copy of the reference checkout where the GPU box can see it.

    python tools/stage_reference.py [/root/reference]

Copies the reference tree (Python sources + its small data files, ~23 MB, no .git) to
``oracle/_ref/MultimodalWordDiscovery/``.  That directory is git-ignored (nothing of the reference enters
the history) but not gpurun-ignored, so it travels with the snapshot exactly like the built ``.so``.  It is
used for two things only, both test / measurement infrastructure:
  * the unchanged-driver proof (tools/unchanged_drivers.py, tests/test_shim_resolution.py): the reference's
    own run_image2phone.py / run_audio.py executed unmodified against the CUDA classes on a B200, and --
    for the image-phone models -- against the reference's own NumPy classes beside them;
  * bench.py's CPU arm (``cpu_baseline.kind == "reference"``): the unmodified NumPy classes timed on the
    box's host cores.
Nothing in the product package reads it.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, 'oracle', '_ref', 'MultimodalWordDiscovery')


def stage(src='/root/reference'):
    if not os.path.isdir(src):
        return None
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    shutil.copytree(src, DEST, ignore=shutil.ignore_patterns('.git', '__pycache__', '*.pyc'))
    return DEST


def find_reference():
    """Reference root for tests / tools: $MWD_REF_ROOT, /root/reference, or the staged copy."""
    for cand in (os.environ.get('MWD_REF_ROOT'), '/root/reference', DEST):
        if cand and os.path.isfile(os.path.join(cand, 'run_image2phone.py')):
            return cand
    return None


if __name__ == '__main__':
    out = stage(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    print('staged:', out)
