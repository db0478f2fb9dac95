"""The image-audio oracle (oracle/image_audio_hmm.py) reproduces the golden vectors of the
unmodified reference class hmm_dnn/image_audio_hmm_word_discoverer.py (CPU only)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, flatten_tables
from oracle import image_audio_hmm as orc


@pytest.mark.parametrize('case', ['short', 'mixed', 'long_floor'])
def test_audio_oracle_matches_reference(case):
    g = np.load(os.path.join(GOLDEN, 'ia_%s.npz' % case))
    fo, ao = g['feat_off'], g['audio_off']
    feats = [g['feats'][fo[i]:fo[i + 1]] for i in range(len(fo) - 1)]
    audio = [g['audio'][ao[i]:ao[i + 1]] for i in range(len(ao) - 1)]
    p = orc.initial_params(feats, int(g['K']), int(g['nPh']), g['WV0'], g['WA0'], lr=float(g['lr']),
                           momentum=float(g['momentum']), phone_probs=g['pp0'] if 'pp0' in g else None)
    lens = [int(v) for v in g['lens']]
    for it in range(int(g['n_iter'])):
        p, info = orc.em_iteration(feats, audio, p)
        np.testing.assert_allclose(info['avg_ll'], g['avg_ll'][it], rtol=1e-10)
        np.testing.assert_allclose(flatten_tables(lens, p['init']), g['init_%d' % it], rtol=1e-9)
        np.testing.assert_allclose(flatten_tables(lens, p['trans']), g['trans_%d' % it], rtol=1e-9)
        np.testing.assert_allclose(p['phone_probs'], g['pp_%d' % it], rtol=1e-9)
        np.testing.assert_allclose(p['WV'], g['WV_%d' % it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(p['WA'], g['WA_%d' % it], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(np.concatenate(info['cC']), g['cC_%d' % it], rtol=1e-9, atol=1e-300)
        # the audio-posterior gradient of the reference is rounding noise (engine_audio.mstep relies on it)
        assert np.abs(info['dWA']).max() < 1e-12
    np.testing.assert_allclose(orc.avg_loglik(feats, audio, p), float(g['final_ll']), rtol=1e-10)
    ali, ics = [], []
    for v, a in zip(feats, audio):
        path, _ = orc.align(v, a, p)
        ali += path
        ics += orc.cluster(v, a, p, path)[0]
    assert np.array_equal(np.array(ali), g['alignment'])
    assert np.array_equal(np.array(ics), g['image_concepts'])


@pytest.mark.parametrize('case', ['short', 'mixed', 'long_unfloored'])
def test_gaussian_audio_oracle_matches_reference(case):
    g = np.load(os.path.join(GOLDEN, 'iag_%s.npz' % case))
    fo, ao = g['feat_off'], g['audio_off']
    feats = [g['feats'][fo[i]:fo[i + 1]] for i in range(len(fo) - 1)]
    audio = [g['audio'][ao[i]:ao[i + 1]] for i in range(len(ao) - 1)]
    p = orc.initial_params_gaussian(feats, int(g['K']), int(g['nPh']), g['musV0'], g['musA0'], width=float(g['width']),
                                    lr=float(g['lr']), momentum=float(g['momentum']),
                                    phone_probs=g['pp0'] if 'pp0' in g else None)
    lens = [int(v) for v in g['lens']]
    for it in range(int(g['n_iter'])):
        p, info = orc.em_iteration_gaussian(feats, audio, p)
        np.testing.assert_allclose(info['avg_ll'], g['avg_ll'][it], rtol=1e-10)
        np.testing.assert_allclose(flatten_tables(lens, p['init']), g['init_%d' % it], rtol=1e-9)
        np.testing.assert_allclose(flatten_tables(lens, p['trans']), g['trans_%d' % it], rtol=1e-9)
        np.testing.assert_allclose(p['phone_probs'], g['pp_%d' % it], rtol=1e-9)
        np.testing.assert_allclose(p['musV'], g['musV_%d' % it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(p['musA'], g['musA_%d' % it], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(np.concatenate(info['cC']), g['cC_%d' % it], rtol=1e-9, atol=1e-300)
        assert np.abs(info['dA']).max() < 1e-10
    ali = []
    for v, a in zip(feats, audio):
        ali += orc.align_gaussian(v, a, p)[0]
    assert np.array_equal(np.array(ali), g['alignment'])
