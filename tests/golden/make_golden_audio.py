#!/usr/bin/env python
"""Golden vectors of ``ImageAudioHMMWordDiscoverer``, produced by running the
UNMODIFIED reference class hmm_dnn/image_audio_hmm_word_discoverer.py from /root/reference.

Run in the build container only:   python tests/golden/make_golden_audio.py

Writes ``tests/golden/ia_<case>.npz``: inputs (region features, audio frame features, injected
WV / WA / phoneProbs) and the reference's per-iteration outputs.  The reference reads only the
first 30 pairs of its input files (:63,:80), so every case has <= 30 pairs.  ``iag_<case>.npz`` are the
same for ``ImageAudioGaussianHMMWordDiscoverer`` (hmm_dnn/image_audio_gaussian_hmm_word_discoverer.py:
RBF posteriors, no EPS floors, no 30-pair cap).
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import flatten_tables, load_ref  # noqa: E402


def synth(rng, N, n_choices, T_lo, T_hi, K, nPh, D, Da):
    cents = rng.standard_normal((K, D))
    acents = rng.standard_normal((nPh, Da))
    feats, audio = [], []
    for _ in range(N):
        n = int(rng.choice(n_choices))
        T = int(rng.integers(T_lo, T_hi + 1))
        v = cents[rng.integers(0, K, n)] + 0.3 * rng.standard_normal((n, D))
        a = acents[rng.integers(0, nPh, T)] + 0.5 * rng.standard_normal((T, Da))
        feats.append(v.astype(np.float32).astype(np.float64))
        audio.append(a.astype(np.float32).astype(np.float64))
    return feats, audio


def run_case(name, feats, audio, K, nPh, n_iter, seed, lr, momentum=0.0, nonuniform=False, w_scale=0.5):
    rng = np.random.default_rng(seed)
    D, Da = feats[0].shape[1], audio[0].shape[1]
    mod = load_ref('hmm_dnn/image_audio_hmm_word_discoverer.py', 'ref_ia_linear')
    with tempfile.TemporaryDirectory() as tmp:
        np.savez(os.path.join(tmp, 'v.npz'), **{'arr_%d' % i: v for i, v in enumerate(feats)})
        np.savez(os.path.join(tmp, 'a.npz'), **{'arr_%d' % i: a for i, a in enumerate(audio)})
        WV0 = w_scale * rng.standard_normal((K, D + 1))
        WA0 = w_scale * rng.standard_normal((nPh, Da + 1))
        np.savez(os.path.join(tmp, 'wv.npz'), weight=WV0[:, :-1], bias=WV0[:, -1])
        np.savez(os.path.join(tmp, 'wa.npz'), weight=WA0[:, :-1], bias=WA0[:, -1])
        cfg = dict(n_words=K, n_phones=nPh, learning_rate=lr, momentum=momentum,
                   image_posterior_weights_file=os.path.join(tmp, 'wv.npz'),
                   audio_posterior_weights_file=os.path.join(tmp, 'wa.npz'))
        pp0 = None
        if nonuniform:
            pp0 = rng.random((K, nPh)) + 0.05
            pp0 /= pp0.sum(1, keepdims=True)
            np.save(os.path.join(tmp, 'pp.npy'), pp0)
            cfg['phone_prob_file'] = os.path.join(tmp, 'pp.npy')
        with contextlib.redirect_stdout(io.StringIO()):
            m = mod.ImageAudioHMMWordDiscoverer(os.path.join(tmp, 'a.npz'), os.path.join(tmp, 'v.npz'), cfg,
                                                modelName=os.path.join(tmp, 'm'))
            m.initializeModel()
        assert len(m.vCorpus) == len(feats) and len(m.aCorpus) == len(audio)
        lens = sorted(m.lenProb)
        out = dict(K=K, nPh=nPh, D=D, Da=Da, n_iter=n_iter, lr=lr, momentum=momentum, lens=np.array(lens),
                   WV0=WV0, WA0=WA0,
                   feat_off=np.cumsum([0] + [v.shape[0] for v in feats]), feats=np.concatenate(feats, axis=0),
                   audio_off=np.cumsum([0] + [a.shape[0] for a in audio]), audio=np.concatenate(audio, axis=0))
        if pp0 is not None:
            out['pp0'] = pp0
        # the oracle is checked against the reference in the same run
        sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        from oracle import image_audio_hmm as orc
        p = orc.initial_params(feats, K, nPh, WV0, WA0, lr=lr, momentum=momentum, phone_probs=pp0)
        lls = []
        for it in range(n_iter):
            with contextlib.redirect_stdout(io.StringIO()):
                m.trainUsingEM(1, warmStart=True, printStatus=True)
            lls.append(np.load(os.path.join(tmp, 'm_likelihoods.npy'))[0])
            out['init_%d' % it] = flatten_tables(lens, m.init)
            out['trans_%d' % it] = flatten_tables(lens, m.trans)
            out['pp_%d' % it] = m.phoneProbs.copy()
            out['WV_%d' % it] = m.WV.copy()
            out['WA_%d' % it] = m.WA.copy()
            out['cC_%d' % it] = np.concatenate(m.conceptCounts, axis=0)
            p, info = orc.em_iteration(feats, audio, p)
            np.testing.assert_allclose(info['avg_ll'], lls[-1], rtol=1e-10)
            np.testing.assert_allclose(flatten_tables(lens, p['init']), out['init_%d' % it], rtol=1e-9)
            np.testing.assert_allclose(flatten_tables(lens, p['trans']), out['trans_%d' % it], rtol=1e-9)
            np.testing.assert_allclose(p['phone_probs'], m.phoneProbs, rtol=1e-9)
            np.testing.assert_allclose(p['WV'], m.WV, rtol=1e-8, atol=1e-12)
            np.testing.assert_allclose(p['WA'], m.WA, rtol=1e-9, atol=1e-13)
            np.testing.assert_allclose(np.concatenate(info['cC']), out['cC_%d' % it], rtol=1e-9, atol=1e-300)
        out['avg_ll'] = np.array(lls)
        with contextlib.redirect_stdout(io.StringIO()):
            out['final_ll'] = m.computeAvgLogLikelihood()
            m.printAlignment(os.path.join(tmp, 'ali'))
        np.testing.assert_allclose(orc.avg_loglik(feats, audio, p), out['final_ll'], rtol=1e-10)
        with open(os.path.join(tmp, 'ali.json')) as f:
            ali = json.load(f)
        assert sorted(ali[0].keys()) == ['align_probs', 'alignment', 'image_concepts', 'index', 'is_phoneme']
        out['alignment'] = np.concatenate([np.array(a['alignment']) for a in ali])
        out['image_concepts'] = np.concatenate([np.array(a['image_concepts']) for a in ali])
        out['align_probs'] = np.concatenate([np.array(a['align_probs']).ravel() for a in ali])
        for ex, (v, a) in enumerate(zip(feats, audio)):
            path, probs = orc.align(v, a, p)
            assert path == ali[ex]['alignment']
            assert orc.cluster(v, a, p, path)[0] == ali[ex]['image_concepts']
            np.testing.assert_allclose(np.array(probs), np.array(ali[ex]['align_probs']), rtol=1e-9)
        out['fwd0'] = m.forward(m.vCorpus[0], m.aCorpus[0])
        out['bwd0'] = m.backward(m.vCorpus[0], m.aCorpus[0])
    np.savez_compressed(os.path.join(HERE, 'ia_%s.npz' % name), **out)
    print('wrote ia_%s.npz  avg_ll=%s  (oracle == reference)' % (name, np.array2string(np.array(lls), precision=6)))


def main():
    rng = np.random.default_rng(20261018)
    # <6 distinct n (Toeplitz pooling off), short utterances: non-floor regime
    f, a = synth(rng, 12, [1, 2, 3, 5], 2, 14, K=13, nPh=9, D=16, Da=12)
    run_case('short', f, a, 13, 9, n_iter=3, seed=1, lr=0.1, nonuniform=True)
    # >= 6 distinct n (Toeplitz pooling on), MSCOCO concept count, default phone-set size, momentum
    f, a = synth(rng, 24, [1, 2, 3, 4, 5, 6, 7], 5, 40, K=65, nPh=42, D=24, Da=20)
    run_case('mixed', f, a, 65, 42, n_iter=3, seed=2, lr=0.05, momentum=0.1)
    # long utterances: raw likelihoods below EPS (floor regime)
    f, a = synth(rng, 8, [3, 5], 90, 130, K=20, nPh=30, D=10, Da=14)
    run_case('long_floor', f, a, 20, 30, n_iter=2, seed=3, lr=0.1)


def run_case_gaussian(name, feats, audio, K, nPh, n_iter, seed, lr, width, momentum=0.0, nonuniform=False):
    """ImageAudioGaussianHMMWordDiscoverer (hmm_dnn/image_audio_gaussian_hmm_word_discoverer.py)."""
    rng = np.random.default_rng(seed)
    D, Da = feats[0].shape[1], audio[0].shape[1]
    mod = load_ref('hmm_dnn/image_audio_gaussian_hmm_word_discoverer.py', 'ref_ia_gauss')
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import image_audio_hmm as orc
    with tempfile.TemporaryDirectory() as tmp:
        np.savez(os.path.join(tmp, 'v.npz'), **{'arr_%d' % i: v for i, v in enumerate(feats)})
        np.savez(os.path.join(tmp, 'a.npz'), **{'arr_%d' % i: a for i, a in enumerate(audio)})
        allv, alla = np.concatenate(feats), np.concatenate(audio)
        musV0 = allv[rng.choice(len(allv), K, replace=len(allv) < K)] + 0.1 * rng.standard_normal((K, D))
        musA0 = alla[rng.choice(len(alla), nPh, replace=len(alla) < nPh)] + 0.1 * rng.standard_normal((nPh, Da))
        np.save(os.path.join(tmp, 'mv.npy'), musV0)
        np.save(os.path.join(tmp, 'ma.npy'), musA0)
        cfg = dict(n_words=K, n_phones=nPh, learning_rate=lr, momentum=momentum, width=width,
                   visual_anchor_file=os.path.join(tmp, 'mv.npy'), audio_anchor_file=os.path.join(tmp, 'ma.npy'))
        pp0 = None
        if nonuniform:
            pp0 = rng.random((K, nPh)) + 0.05
            pp0 /= pp0.sum(1, keepdims=True)
            np.save(os.path.join(tmp, 'pp.npy'), pp0)
            cfg['phone_prob_file'] = os.path.join(tmp, 'pp.npy')
        with contextlib.redirect_stdout(io.StringIO()):
            m = mod.ImageAudioGaussianHMMWordDiscoverer(os.path.join(tmp, 'a.npz'), os.path.join(tmp, 'v.npz'), cfg,
                                                        modelName=os.path.join(tmp, 'm'))
            m.initializeModel()
        lens = sorted(m.lenProb)
        out = dict(K=K, nPh=nPh, D=D, Da=Da, n_iter=n_iter, lr=lr, momentum=momentum, width=width,
                   lens=np.array(lens), musV0=musV0, musA0=musA0,
                   feat_off=np.cumsum([0] + [v.shape[0] for v in feats]), feats=np.concatenate(feats, axis=0),
                   audio_off=np.cumsum([0] + [a.shape[0] for a in audio]), audio=np.concatenate(audio, axis=0))
        if pp0 is not None:
            out['pp0'] = pp0
        p = orc.initial_params_gaussian(feats, K, nPh, musV0, musA0, width=width, lr=lr, momentum=momentum,
                                        phone_probs=pp0)
        lls = []
        for it in range(n_iter):
            with contextlib.redirect_stdout(io.StringIO()):
                m.trainUsingEM(1, warmStart=True, printStatus=True)
            lls.append(np.load(os.path.join(tmp, 'm_likelihoods.npy'))[0])
            out['init_%d' % it] = flatten_tables(lens, m.init)
            out['trans_%d' % it] = flatten_tables(lens, m.trans)
            out['pp_%d' % it] = m.phoneProbs.copy()
            out['musV_%d' % it] = m.musV.copy()
            out['musA_%d' % it] = m.musA.copy()
            out['cC_%d' % it] = np.concatenate(m.conceptCounts, axis=0)
            p, info = orc.em_iteration_gaussian(feats, audio, p)
            np.testing.assert_allclose(info['avg_ll'], lls[-1], rtol=1e-10)
            np.testing.assert_allclose(flatten_tables(lens, p['init']), out['init_%d' % it], rtol=1e-9)
            np.testing.assert_allclose(flatten_tables(lens, p['trans']), out['trans_%d' % it], rtol=1e-9)
            np.testing.assert_allclose(p['phone_probs'], m.phoneProbs, rtol=1e-9)
            np.testing.assert_allclose(p['musV'], m.musV, rtol=1e-8, atol=1e-12)
            np.testing.assert_allclose(p['musA'], m.musA, rtol=1e-9, atol=1e-13)
            np.testing.assert_allclose(np.concatenate(info['cC']), out['cC_%d' % it], rtol=1e-9, atol=1e-300)
            assert np.abs(info['dA']).max() < 1e-10, np.abs(info['dA']).max()
        out['avg_ll'] = np.array(lls)
        with contextlib.redirect_stdout(io.StringIO()):
            out['final_ll'] = m.computeAvgLogLikelihood()
            m.printAlignment(os.path.join(tmp, 'ali'))
        with open(os.path.join(tmp, 'ali.json')) as f:
            ali = json.load(f)
        out['ali_keys'] = np.array(sorted(ali[0].keys()))
        out['alignment'] = np.concatenate([np.array(a['alignment']) for a in ali])
        out['image_concepts'] = np.concatenate([np.array(a['image_concepts']) for a in ali])
        out['phone_clusters'] = np.concatenate([np.array(a['phone_clusters']) for a in ali])
        out['concept_alignment'] = np.concatenate([np.array(a['concept_alignment']) for a in ali])
        out['align_probs'] = np.concatenate([np.array(a['align_probs']).ravel() for a in ali])
        out['concept_probs'] = np.concatenate([np.array(a['concept_probs']).ravel() for a in ali])
        for ex, (v, a) in enumerate(zip(feats, audio)):
            path, probs = orc.align_gaussian(v, a, p)
            assert path == ali[ex]['alignment']
            assert orc.cluster_gaussian(v, a, p, path)[0] == ali[ex]['image_concepts']
        out['fwd0'] = m.forward(m.vCorpus[0], m.aCorpus[0])
        out['bwd0'] = m.backward(m.vCorpus[0], m.aCorpus[0])
    np.savez_compressed(os.path.join(HERE, 'iag_%s.npz' % name), **out)
    print('wrote iag_%s.npz  avg_ll=%s  (oracle == reference)' % (name, np.array2string(np.array(lls), precision=6)))


def main_gaussian():
    rng = np.random.default_rng(20261019)
    f, a = synth(rng, 12, [1, 2, 3, 5], 2, 14, K=13, nPh=9, D=16, Da=12)
    run_case_gaussian('short', f, a, 13, 9, n_iter=3, seed=11, lr=0.1, width=8.0, nonuniform=True)
    # >= 6 distinct n (Toeplitz pooling), 40 pairs (this class has no 30-pair cap), momentum
    f, a = synth(rng, 40, [1, 2, 3, 4, 5, 6, 7], 5, 40, K=65, nPh=42, D=24, Da=20)
    run_case_gaussian('mixed', f, a, 65, 42, n_iter=3, seed=12, lr=0.05, width=12.0, momentum=0.1)
    # long utterances: raw likelihoods ~1e-150, far below EPS -- the un-floored normalisers still work
    f, a = synth(rng, 8, [3, 5], 90, 130, K=20, nPh=30, D=10, Da=14)
    run_case_gaussian('long_unfloored', f, a, 20, 30, n_iter=2, seed=13, lr=0.1, width=6.0)


if __name__ == '__main__':
    main()
    main_gaussian()
