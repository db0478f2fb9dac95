"""GPU parity of ``ImageAudioHMMWordDiscoverer``: the CUDA-backed class mirror vs
golden vectors of the unmodified reference class (tests/golden/make_golden_audio.py) -- log-likelihood
and tables within 1e-9 relative (north star: 1e-5), Viterbi alignments / argmax concepts bit-exact."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, flatten_tables

pytestmark = pytest.mark.gpu

RTOL = 1e-9
CASES = ['short', 'mixed', 'long_floor']


def _load(case):
    g = dict(np.load(os.path.join(GOLDEN, 'ia_%s.npz' % case)))
    fo, ao = g['feat_off'], g['audio_off']
    g['feats_list'] = [g['feats'][fo[i]:fo[i + 1]] for i in range(len(fo) - 1)]
    g['audio_list'] = [g['audio'][ao[i]:ao[i + 1]] for i in range(len(ao) - 1)]
    return g


def _model(g, tmp_path):
    from multimodalworddiscovery_b200.hmm_dnn.image_audio_hmm_word_discoverer import ImageAudioHMMWordDiscoverer
    tmp = str(tmp_path)
    np.savez(os.path.join(tmp, 'v.npz'), **{'arr_%d' % i: v for i, v in enumerate(g['feats_list'])})
    np.savez(os.path.join(tmp, 'a.npz'), **{'arr_%d' % i: a for i, a in enumerate(g['audio_list'])})
    np.savez(os.path.join(tmp, 'wv.npz'), weight=g['WV0'][:, :-1], bias=g['WV0'][:, -1])
    np.savez(os.path.join(tmp, 'wa.npz'), weight=g['WA0'][:, :-1], bias=g['WA0'][:, -1])
    cfg = dict(n_words=int(g['K']), n_phones=int(g['nPh']), learning_rate=float(g['lr']), momentum=float(g['momentum']),
               image_posterior_weights_file=os.path.join(tmp, 'wv.npz'),
               audio_posterior_weights_file=os.path.join(tmp, 'wa.npz'), feature_dtype='float64')
    if 'pp0' in g:
        np.save(os.path.join(tmp, 'pp.npy'), g['pp0'])
        cfg['phone_prob_file'] = os.path.join(tmp, 'pp.npy')
    m = ImageAudioHMMWordDiscoverer(os.path.join(tmp, 'a.npz'), os.path.join(tmp, 'v.npz'), cfg,
                                    modelName=os.path.join(tmp, 'm'))
    m.initializeModel()
    return m


@pytest.mark.parametrize('case', CASES)
def test_image_audio_class_matches_reference(case, tmp_path):
    g = _load(case)
    m = _model(g, tmp_path)
    lens = [int(v) for v in g['lens']]
    assert sorted(m.lenProb) == lens
    for it in range(int(g['n_iter'])):
        m.trainUsingEM(1, warmStart=True, printStatus=True)
        ll = np.load(os.path.join(str(tmp_path), 'm_likelihoods.npy'))[0]
        np.testing.assert_allclose(ll, g['avg_ll'][it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(lens, m.init), g['init_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(lens, m.trans), g['trans_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(m.phoneProbs, g['pp_%d' % it], rtol=RTOL, atol=0)
        np.testing.assert_allclose(m.WV, g['WV_%d' % it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(m.WA, g['WA_%d' % it], rtol=RTOL, atol=1e-15)
        np.testing.assert_allclose(np.concatenate(m.conceptCounts, axis=0), g['cC_%d' % it], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(m.computeAvgLogLikelihood(), float(g['final_ll']), rtol=RTOL)
    # printAlignment: bit-exact integers, reference key set
    m.printAlignment(os.path.join(str(tmp_path), 'ali'))
    with open(os.path.join(str(tmp_path), 'ali.json')) as f:
        ali = json.load(f)
    assert sorted(ali[0].keys()) == ['align_probs', 'alignment', 'image_concepts', 'index', 'is_phoneme']
    assert np.array_equal(np.concatenate([a['alignment'] for a in ali]), g['alignment'])
    assert np.array_equal(np.concatenate([a['image_concepts'] for a in ali]), g['image_concepts'])
    np.testing.assert_allclose(np.concatenate([np.array(a['align_probs']).ravel() for a in ali]), g['align_probs'],
                               rtol=1e-8)
    # single-pair API
    v0, a0 = m.vCorpus[0], m.aCorpus[0]
    np.testing.assert_allclose(m.forward(v0, a0), g['fwd0'], rtol=RTOL)
    np.testing.assert_allclose(m.backward(v0, a0), g['bwd0'], rtol=RTOL)
    path, probs = m.align(a0, v0)
    assert path == ali[0]['alignment']
    assert m.cluster(a0, v0, path)[0] == ali[0]['image_concepts']
    np.testing.assert_allclose(m.softmaxLayerA(a0).sum(1), 1.0, rtol=1e-12)
    # printModel writes the reference's file set
    m.printModel(os.path.join(str(tmp_path), 'pm'))
    for suffix in ('_initialprobs.txt', '_transitionprobs.txt', '_phoneprobs.npy', '_phone2idx.json',
                   '_visual_posterior_weights.npy', '_audio_posterior_weights.npy'):
        assert os.path.exists(os.path.join(str(tmp_path), 'pm' + suffix))


def test_image_audio_matches_oracle_beyond_30_pairs(tmp_path):
    """The reference reads only 30 pairs; with the cap lifted (``pair_limit``) the CUDA path is
    checked against the oracle on a 64-pair corpus at the MSCOCO concept count."""
    from oracle import image_audio_hmm as orc
    from multimodalworddiscovery_b200.engine_audio import IKAudioEngine, pack_audio_pairs
    rng = np.random.default_rng(5)
    K, nPh, D, Da = 65, 42, 32, 24
    feats, audio = [], []
    for _ in range(64):
        n, T = int(rng.integers(1, 9)), int(rng.integers(1, 45))
        feats.append(rng.standard_normal((n, D)).astype(np.float32).astype(np.float64))
        audio.append(rng.standard_normal((T, Da)).astype(np.float32).astype(np.float64))
    pp0 = rng.random((K, nPh)) + 0.05
    pp0 /= pp0.sum(1, keepdims=True)
    p = orc.initial_params(feats, K, nPh, 0.3 * rng.standard_normal((K, D + 1)), 0.3 * rng.standard_normal((nPh, Da + 1)),
                           lr=0.1, momentum=0.05, phone_probs=pp0)
    pk, aud = pack_audio_pairs(feats, audio, feat_dtype=np.float64)
    eng = IKAudioEngine(pk, aud, K, nPh)
    eng.set_params(p['init'], p['trans'], p['phone_probs'], p['WV'])
    eng.set_audio_param(p['WA'])
    lens = sorted(p['init'])
    for it in range(2):
        p, info = orc.em_iteration(feats, audio, p)
        ll = eng.em_iteration(0.1, 0.05)
        np.testing.assert_allclose(float(ll) / len(feats), info['avg_ll'], rtol=RTOL)
        init, trans, pp, WV = eng.get_params()
        np.testing.assert_allclose(flatten_tables(lens, init), flatten_tables(lens, p['init']), rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(lens, trans), flatten_tables(lens, p['trans']), rtol=RTOL)
        np.testing.assert_allclose(pp, p['phone_probs'], rtol=RTOL)
        np.testing.assert_allclose(WV, p['WV'], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(eng.get_audio_param(), p['WA'], rtol=RTOL, atol=1e-15)


@pytest.mark.parametrize('case', ['short', 'mixed', 'long_unfloored'])
def test_image_audio_gaussian_class_matches_reference(case, tmp_path):
    """ImageAudioGaussianHMMWordDiscoverer: RBF posteriors on both sides, NO EPS floors (the long case has
    raw likelihoods ~1e-165, far below EPS, and still normalised counts)."""
    from multimodalworddiscovery_b200.hmm_dnn.image_audio_gaussian_hmm_word_discoverer import \
        ImageAudioGaussianHMMWordDiscoverer
    g = dict(np.load(os.path.join(GOLDEN, 'iag_%s.npz' % case)))
    fo, ao = g['feat_off'], g['audio_off']
    feats = [g['feats'][fo[i]:fo[i + 1]] for i in range(len(fo) - 1)]
    audio = [g['audio'][ao[i]:ao[i + 1]] for i in range(len(ao) - 1)]
    tmp = str(tmp_path)
    np.savez(os.path.join(tmp, 'v.npz'), **{'arr_%d' % i: v for i, v in enumerate(feats)})
    np.savez(os.path.join(tmp, 'a.npz'), **{'arr_%d' % i: a for i, a in enumerate(audio)})
    np.save(os.path.join(tmp, 'mv.npy'), g['musV0'])
    np.save(os.path.join(tmp, 'ma.npy'), g['musA0'])
    cfg = dict(n_words=int(g['K']), n_phones=int(g['nPh']), learning_rate=float(g['lr']), momentum=float(g['momentum']),
               width=float(g['width']), visual_anchor_file=os.path.join(tmp, 'mv.npy'),
               audio_anchor_file=os.path.join(tmp, 'ma.npy'), feature_dtype='float64')
    if 'pp0' in g:
        np.save(os.path.join(tmp, 'pp.npy'), g['pp0'])
        cfg['phone_prob_file'] = os.path.join(tmp, 'pp.npy')
    m = ImageAudioGaussianHMMWordDiscoverer(os.path.join(tmp, 'a.npz'), os.path.join(tmp, 'v.npz'), cfg,
                                            modelName=os.path.join(tmp, 'm'))
    assert len(m.vCorpus) == len(feats)
    m.initializeModel()
    assert os.path.exists(os.path.join(tmp, 'm.json'))              # printUnimodalCluster (:150)
    lens = [int(v) for v in g['lens']]
    for it in range(int(g['n_iter'])):
        m.trainUsingEM(1, warmStart=True, printStatus=True)
        ll = np.load(os.path.join(tmp, 'm_likelihoods.npy'))[0]
        np.testing.assert_allclose(ll, g['avg_ll'][it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(lens, m.init), g['init_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(lens, m.trans), g['trans_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(m.phoneProbs, g['pp_%d' % it], rtol=RTOL, atol=0)
        np.testing.assert_allclose(m.musV, g['musV_%d' % it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(m.musA, g['musA_%d' % it], rtol=RTOL, atol=1e-15)
        np.testing.assert_allclose(np.concatenate(m.conceptCounts, axis=0), g['cC_%d' % it], rtol=RTOL, atol=1e-300)
    cpc = m.conceptPhoneCounts
    assert cpc[0].shape == (len(audio[0]), int(g['K']), int(g['nPh']))
    np.testing.assert_allclose(cpc[0].sum((1, 2)), 1.0, rtol=1e-12)
    np.testing.assert_allclose(m.computeAvgLogLikelihood(), float(g['final_ll']), rtol=RTOL)
    m.printAlignment(os.path.join(tmp, 'ali'))
    with open(os.path.join(tmp, 'ali.json')) as f:
        ali = json.load(f)
    assert sorted(ali[0].keys()) == [str(k) for k in g['ali_keys']]
    for key in ('alignment', 'image_concepts', 'phone_clusters', 'concept_alignment'):
        assert np.array_equal(np.concatenate([a[key] for a in ali]), g[key]), key
    np.testing.assert_allclose(np.concatenate([np.array(a['align_probs']).ravel() for a in ali]), g['align_probs'],
                               rtol=1e-8)
    np.testing.assert_allclose(np.concatenate([np.array(a['concept_probs']).ravel() for a in ali]),
                               g['concept_probs'], rtol=RTOL, atol=1e-300)
    v0, a0 = m.vCorpus[0], m.aCorpus[0]
    np.testing.assert_allclose(m.forward(v0, a0), g['fwd0'], rtol=RTOL)
    np.testing.assert_allclose(m.backward(v0, a0), g['bwd0'], rtol=RTOL)
    path, _ = m.align(a0, v0)
    assert path == ali[0]['alignment']
    assert m.cluster(a0, v0, path)[0] == ali[0]['image_concepts']
