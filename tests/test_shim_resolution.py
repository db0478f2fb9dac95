"""The shim packages shadow the reference's namespace packages: with the reference
directory FIRST on sys.path (as when its driver script runs), `hmm_dnn.*` / `hmm.*` still resolve
to the CUDA-backed classes, and the reference's driver-side imports (utils/) load with the stubs.
Runs only where /root/reference exists (the build container)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present')
def test_driver_star_imports_resolve_to_b200_classes():
    code = r'''
import sys
sys.path.insert(0, %r)                      # the driver's own directory comes first, like `python run_image2phone.py`
from hmm_dnn.image_phone_hmm_word_discoverer import *
from hmm_dnn.image_phone_gaussian_hmm_word_discoverer import *
from hmm.hmm_word_discoverer import *
from hmm.audio_hmm_word_discoverer import *
from utils.clusteval import *               # reference module; needs the nltk / matplotlib stubs
from utils.postprocess import *
for cls in (ImagePhoneHMMWordDiscoverer, ImagePhoneGaussianHMMWordDiscoverer, HMMWordDiscoverer, AudioHMMWordDiscoverer):
    assert cls.__module__.startswith('multimodalworddiscovery_b200.'), cls.__module__
assert np.__name__ == 'numpy' and json.__name__ == 'json'   # names the drivers rely on (run_image2phone.py:132,137)
print('OK')
''' % REF
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 'shim'), ROOT]))
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, cwd='/tmp')
    assert out.returncode == 0 and 'OK' in out.stdout, out.stderr[-2000:]
