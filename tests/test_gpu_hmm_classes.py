"""hmm/ class mirrors (HMMWordDiscoverer, AudioHMMWordDiscoverer) against the reference goldens."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

from helpers import HMM_CASES, flatten_tables, load_hmm

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _write_corpus(path, g):
    tv, fv = g['tgt_vocab'], g['src_vocab']
    with open(path, 'w') as f:
        for e, s in zip(g['tgt_list'], g['src_list']):
            words = [str(tv[i]) for i in e]
            if g['kind'] == 'log':
                assert words[0] == 'NULL'
                words = words[1:]
            f.write(' '.join(words) + '\n' + ' '.join(str(fv[i]) for i in s) + '\n\n')


@pytest.mark.parametrize('case', HMM_CASES)
def test_hmm_class_matches_reference(case, tmp_path):
    g = load_hmm(case)
    path = str(tmp_path / 'corpus.txt')
    _write_corpus(path, g)
    lens = [int(m) for m in g['lens']]
    with contextlib.redirect_stdout(io.StringIO()):
        if g['kind'] == 'log':
            from multimodalworddiscovery_b200.hmm.audio_hmm_word_discoverer import AudioHMMWordDiscoverer
            m = AudioHMMWordDiscoverer(path, modelName=str(tmp_path / 'm'))
        else:
            from multimodalworddiscovery_b200.hmm.hmm_word_discoverer import HMMWordDiscoverer
            m = HMMWordDiscoverer(path, modelName=str(tmp_path / 'm'))
    assert sorted(m.lenProb) == lens
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m.trainUsingEM(g['n_iter'])
    lls = [float(ln.split(':')[-1]) for ln in buf.getvalue().split('\n') if 'Average Log Likelihood' in ln]
    np.testing.assert_allclose(lls, g['avg_ll'], rtol=RTOL)
    it = g['n_iter'] - 1
    np.testing.assert_allclose(flatten_tables(lens, m.init), g['init_%d' % it], rtol=RTOL)
    np.testing.assert_allclose(flatten_tables(lens, m.trans), g['trans_%d' % it], rtol=RTOL)
    # dict-of-dict view of obs
    tv, fv = g['tgt_vocab'], g['src_vocab']
    ref = g['obs_%d' % it]
    n_present = 0
    for r in range(ref.shape[0]):
        for c in range(ref.shape[1]):
            if not np.isnan(ref[r, c]):
                n_present += 1
                assert m.obs[str(tv[r])][str(fv[c])] == pytest.approx(ref[r, c], rel=RTOL)
    assert n_present == sum(len(v) for v in m.obs.values())
    np.testing.assert_allclose(m.computeAvgLogLikelihood(), float(g['final_ll']), rtol=RTOL)
    with contextlib.redirect_stdout(io.StringIO()):
        m.printAlignment(str(tmp_path / 'ali'))
        m.printModel(str(tmp_path / 'model'))
    ali = json.load(open(str(tmp_path / 'ali.json')))
    assert np.array_equal(np.concatenate([a['alignment'] for a in ali]), g['alignment'])
    np.testing.assert_allclose(np.concatenate([np.array(a['align_probs']).ravel() for a in ali]),
                               g['align_probs'], rtol=1e-8)
    assert ('is_audio' in ali[0]) == (g['kind'] == 'log')
    assert os.path.exists(str(tmp_path / 'model_observationprobs.txt'))
    e0, f0 = m.tCorpus[0], m.fCorpus[0]
    np.testing.assert_allclose(m.forward(e0, f0), g['fwd0'], rtol=RTOL)
    np.testing.assert_allclose(m.backward(e0, f0), g['bwd0'], rtol=RTOL)
    path0, probs0 = m.align(f0, e0)
    assert path0 == g['alignment'][:len(f0)].tolist()
