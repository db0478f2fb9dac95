"""The shim packages shadow the reference's namespace packages: with the reference
directory FIRST on sys.path (as when its driver script runs), `hmm_dnn.*` / `hmm.*` still resolve
to the CUDA-backed classes, and the reference's driver-side imports (utils/) load with the stubs.
The reference checkout is looked up in $MWD_REF_ROOT, /root/reference (build container) and the untracked
copy staged by tools/stage_reference.py under oracle/_ref/ (which travels to the GPU box), so the
unchanged-driver test completes on a B200 (-m gpu) and reaches the no-CPU-fallback error without one."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from stage_reference import find_reference  # noqa: E402

REF = find_reference() or '/nonexistent'
PYPATH = os.pathsep.join([os.path.join(ROOT, 'shim'), os.path.join(ROOT, 'shim_stubs'), ROOT])


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present')
def test_driver_star_imports_resolve_to_b200_classes():
    code = r'''
import sys
sys.path.insert(0, %r)                      # the driver's own directory comes first, like `python run_image2phone.py`
from hmm_dnn.image_phone_hmm_word_discoverer import *
from hmm_dnn.image_phone_hmm_dnn_word_discoverer import *      # run_image2phone.py:2
from hmm_dnn.image_phone_gaussian_hmm_word_discoverer import *
from hmm_dnn.image_audio_hmm_word_discoverer import *          # run_image2audio.py:9
from hmm_dnn.image_audio_gaussian_hmm_word_discoverer import * # run_image2audio.py:10
from hmm_dnn.image_phone_bhmm_word_discoverer import *          # reference file (not mirrored)
from hmm.hmm_word_discoverer import *
from hmm.audio_segembed_hmm_word_discoverer import *
from hmm.audio_hmm_word_discoverer import *
from utils.clusteval import *               # reference module; needs the nltk / matplotlib stubs
from utils.postprocess import *
for cls in (ImagePhoneHMMWordDiscoverer, ImagePhoneGaussianHMMWordDiscoverer, HMMWordDiscoverer, AudioHMMWordDiscoverer,
            SegEmbedHMMWordDiscoverer, ImagePhoneHMMDNNWordDiscoverer, ImageAudioHMMWordDiscoverer,
            ImageAudioGaussianHMMWordDiscoverer):
    assert cls.__module__.startswith('multimodalworddiscovery_b200.'), cls.__module__
assert ImagePhoneBigramHMMWordDiscoverer.__module__ == 'hmm_dnn.image_phone_bhmm_word_discoverer'
assert np.__name__ == 'numpy' and json.__name__ == 'json'   # names the drivers rely on (run_image2phone.py:132,137)
print('OK')
''' % REF
    env = dict(os.environ, PYTHONPATH=PYPATH)
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, cwd='/tmp')
    assert out.returncode == 0 and 'OK' in out.stdout, out.stderr[-2000:]


def _run_unchanged_image2phone(tmp_path, model_type='linear'):
    """Run the reference's run_image2phone.py UNCHANGED from a scratch directory laid out as the
    driver expects (CWD-relative data/mscoco/..., an existing hmm_dnn/exp/).  It must construct the
    B200 class, print the reference's corpus summary and enter trainUsingEM; without a GPU the
    engine then fails loudly (no CPU fallback), with a GPU it completes and writes the alignment."""
    import numpy as np
    import torch
    rng = np.random.default_rng(0)
    data = tmp_path / 'data' / 'mscoco'
    data.mkdir(parents=True)
    (tmp_path / 'hmm_dnn' / 'exp').mkdir(parents=True)
    feats, caps = {}, []
    for i in range(6):
        feats['arr_%d' % i] = rng.standard_normal((int(rng.integers(1, 4)), 16)).astype(np.float32)
        caps.append(' '.join('p%d' % p for p in rng.integers(0, 7, int(rng.integers(3, 9)))))
    np.savez(str(data / 'mscoco2k_res34_embed512dim.npz'), **feats)
    (data / 'mscoco2k_phone_captions.txt').write_text('\n'.join(caps) + '\n')
    env = dict(os.environ, PYTHONPATH=PYPATH)
    out = subprocess.run([sys.executable, os.path.join(REF, 'run_image2phone.py'), '--dataset', 'mscoco2k',
                          '--feat_type', 'res34', '--model_type', model_type, '--lr', '0.01', '--hidden_dim', '8'],
                         env=env, capture_output=True, text=True, cwd=str(tmp_path))
    assert 'Start training the model ...' in out.stdout, out.stderr[-1500:]
    assert '----- Corpus Summary -----' in out.stdout and 'Number of examples:  6' in out.stdout
    return out


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present')
def test_unchanged_driver_reaches_the_cuda_path(tmp_path):
    """Without a GPU the unchanged driver constructs the B200 class, prints the corpus summary, enters
    trainUsingEM and fails loudly (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present: covered by test_unchanged_driver_completes_on_gpu')
    out = _run_unchanged_image2phone(tmp_path)
    assert out.returncode != 0 and 'MwdError' in out.stderr and 'no CPU fallback' in out.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present (tools/stage_reference.py)')
@pytest.mark.parametrize('model_type', ['linear', 'two-layer'])
def test_unchanged_driver_completes_on_gpu(tmp_path, model_type):
    """run_image2phone.py, unmodified, to completion on the CUDA classes: 20 EM iterations + printAlignment.
    (tools/unchanged_drivers.py runs the larger version, all three model types + run_audio.py, next to the
    reference's own classes; its summary is committed under profiles/.)"""
    import json
    out = _run_unchanged_image2phone(tmp_path, model_type)
    assert out.returncode == 0, out.stderr[-1500:]
    assert 'to finish decoding' in out.stdout
    exp = [d for d in (tmp_path / 'hmm_dnn' / 'exp').iterdir()][0]
    ali = json.load(open(str(exp / 'image_phone_alignment.json')))
    assert len(ali) == 6 and all(len(a['alignment']) >= 3 for a in ali)
