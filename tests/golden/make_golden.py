#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference classes from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes ``tests/golden/ik_<case>.npz`` (hmm_dnn linear / gaussian classes) and
``tests/golden/hmm_<case>.npz`` (hmm/ classes).  Every case stores its *inputs* (features, phone
ids, injected parameters) next to the reference's per-iteration outputs, so the tests need nothing
but the .npz files.  All randomness comes from seeded ``np.random.default_rng`` generators; the
reference's own global-RNG draws are bypassed by injecting W / mus / obs through files.
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import tempfile

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref(relpath, alias):
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


def synth_ik_corpus(rng, N, n_choices, T_lo, T_hi, K, P, D, feat_scale=1.0, zipf=True):
    """MSCOCO-shaped synthetic pairs: gaussian-cluster region features + zipfian phones."""
    centroids = feat_scale * rng.standard_normal((K, D))
    pw = 1.0 / np.arange(1, P + 1) ** 1.2 if zipf else np.ones(P)
    pw /= pw.sum()
    feats, phones = [], []
    for _ in range(N):
        n = int(rng.choice(n_choices))
        T = int(rng.integers(T_lo, T_hi + 1))
        c = rng.integers(0, K, size=n)
        v = centroids[c] + 0.3 * rng.standard_normal((n, D))
        feats.append(v.astype(np.float32).astype(np.float64))   # fp32-representable
        phones.append(rng.choice(P, size=T, p=pw).astype(np.int64))
    # every phone id must occur, otherwise the reference's phone2idx (first-seen order) is smaller
    seen = np.unique(np.concatenate(phones))
    missing = [p for p in range(P) if p not in seen]
    if missing:
        phones[0] = np.concatenate([phones[0], np.array(missing, dtype=np.int64)])
    return feats, phones


def write_ik_files(tmp, feats, phones):
    """Reference on-disk formats (image_phone_hmm_word_discoverer.py:51-53,78-88).  Phone ids are
    renamed so that the reference's first-seen ``phone2idx`` equals our integer ids."""
    order = []
    for x in phones:
        for p in x:
            if int(p) not in order:
                order.append(int(p))
    remap = {p: i for i, p in enumerate(order)}
    phones2 = [np.array([remap[int(p)] for p in x], dtype=np.int64) for x in phones]
    with open(os.path.join(tmp, 'caps.txt'), 'w') as f:
        for x in phones2:
            f.write(' '.join('p%d' % p for p in x) + '\n')
    np.savez(os.path.join(tmp, 'feats.npz'), **{'arr_%d' % i: v for i, v in enumerate(feats)})
    return phones2


def flatten_tables(lens, tabs):
    return np.concatenate([np.asarray(tabs[m], dtype=np.float64).ravel() for m in lens])


def run_ik_case(name, kind, feats, phones, K, n_iter, seed, lr, momentum=0.0, width=1.0,
                w_scale=0.5, nonuniform_obs=False, hidden_dim=0):
    rng = np.random.default_rng(seed)
    D = feats[0].shape[1]
    with tempfile.TemporaryDirectory() as tmp:
        phones = write_ik_files(tmp, feats, phones)
        P = int(max(x.max() for x in phones)) + 1
        cfg = dict(has_null=False, n_words=K, learning_rate=lr, momentum=momentum, width=width)
        obs0 = None
        if nonuniform_obs:
            obs0 = rng.random((K, P)) + 0.05
            obs0 /= obs0.sum(1, keepdims=True)
            np.save(os.path.join(tmp, 'obs.npy'), obs0)
        if kind == 'linear':
            mod = load_ref('hmm_dnn/image_phone_hmm_word_discoverer.py', 'ref_ik_linear')
            W0 = w_scale * rng.standard_normal((K, D + 1))
            np.savez(os.path.join(tmp, 'w.npz'), weight=W0[:, :-1], bias=W0[:, -1])
            cfg['image_posterior_weights_file'] = os.path.join(tmp, 'w.npz')
            with contextlib.redirect_stdout(io.StringIO()):
                m = mod.ImagePhoneHMMWordDiscoverer(
                    os.path.join(tmp, 'caps.txt'), os.path.join(tmp, 'feats.npz'), cfg,
                    obsProbFile=os.path.join(tmp, 'obs.npy') if nonuniform_obs else None,
                    modelName=os.path.join(tmp, 'm'))
            P0 = W0
        elif kind == 'two-layer':
            mod = load_ref('hmm_dnn/image_phone_hmm_dnn_word_discoverer.py', 'ref_ik_twolayer')
            H = hidden_dim
            V0 = w_scale * rng.standard_normal((H, D + 1))
            W0 = w_scale * rng.uniform(-1., 1., size=(K, H + 1))
            np.savez(os.path.join(tmp, 'w.npz'), arr_0=V0[:, :-1], arr_1=V0[:, -1], arr_2=W0[:, :-1], arr_3=W0[:, -1])
            cfg['image_posterior_weights_file'] = os.path.join(tmp, 'w.npz')
            cfg['hidden_dim'] = H
            with contextlib.redirect_stdout(io.StringIO()):
                m = mod.ImagePhoneHMMDNNWordDiscoverer(
                    os.path.join(tmp, 'caps.txt'), os.path.join(tmp, 'feats.npz'), cfg,
                    obsProbFile=os.path.join(tmp, 'obs.npy') if nonuniform_obs else None,
                    modelName=os.path.join(tmp, 'm'))
            P0 = W0
        else:
            mod = load_ref('hmm_dnn/image_phone_gaussian_hmm_word_discoverer.py', 'ref_ik_gauss')
            mus0 = w_scale * rng.standard_normal((K, D))
            np.save(os.path.join(tmp, 'mus.npy'), mus0)
            cfg['visual_anchor_file'] = os.path.join(tmp, 'mus.npy')
            if nonuniform_obs:
                cfg['obs_prob_file'] = os.path.join(tmp, 'obs.npy')
            with contextlib.redirect_stdout(io.StringIO()):
                m = mod.ImagePhoneGaussianHMMWordDiscoverer(
                    os.path.join(tmp, 'caps.txt'), os.path.join(tmp, 'feats.npz'), cfg,
                    modelName=os.path.join(tmp, 'm'))
            assert len(m.vCorpus) == len(feats), 'gaussian reference truncates to 30 pairs'
            P0 = mus0
        assert m.audioFeatDim == P
        out = dict(kind=kind, K=K, P=P, D=D, n_iter=n_iter, lr=lr, momentum=momentum, width=width,
                   param0=P0,
                   feat_off=np.cumsum([0] + [v.shape[0] for v in feats]),
                   feats=np.concatenate(feats, axis=0),
                   phone_off=np.cumsum([0] + [len(x) for x in phones]),
                   phones=np.concatenate(phones))
        if obs0 is not None:
            out['obs0'] = obs0
        if kind == 'two-layer':
            out['hidden0'] = V0
            out['hidden_dim'] = hidden_dim
        with contextlib.redirect_stdout(io.StringIO()):
            m.initializeModel()
        lens = sorted(m.lenProb)
        out['lens'] = np.array(lens)
        # per-iteration tables: run one epoch at a time with warmStart (the reference's lr decay
        # `(epoch+1) % 10` therefore never triggers; the tests replay the same protocol)
        lls = []
        for it in range(n_iter):
            with contextlib.redirect_stdout(io.StringIO()):
                m.trainUsingEM(1, warmStart=True, printStatus=True)
            lls.append(np.load(os.path.join(tmp, 'm_likelihoods.npy'))[0])
            out['init_%d' % it] = flatten_tables(lens, m.init)
            out['trans_%d' % it] = flatten_tables(lens, m.trans)
            out['obs_%d' % it] = m.obs.copy()
            out['param_%d' % it] = (m.mus if kind == 'gaussian' else m.W).copy()
            if kind == 'two-layer':
                out['hidden_%d' % it] = m.V.copy()       # the class keeps neither conceptCounts nor conceptCountsA
            else:
                out['cC_%d' % it] = np.concatenate(m.conceptCounts, axis=0)
                out['cA_%d' % it] = np.concatenate(m.conceptCountsA, axis=0)
        out['avg_ll'] = np.array(lls)
        with contextlib.redirect_stdout(io.StringIO()):
            out['final_ll'] = m.computeAvgLogLikelihood()
            m.printAlignment(os.path.join(tmp, 'ali'))
        with open(os.path.join(tmp, 'ali.json')) as f:
            ali = json.load(f)
        out['alignment'] = np.concatenate([np.array(a['alignment']) for a in ali])
        out['image_concepts'] = np.concatenate([np.array(a['image_concepts']) for a in ali])
        if kind == 'two-layer':
            out['cluster_probs'] = np.concatenate([np.array(a['cluster_probs']).ravel() for a in ali])
        else:
            out['concept_alignment'] = np.concatenate([np.array(a['concept_alignment']) for a in ali])
        out['align_probs'] = np.concatenate([np.array(a['align_probs']).ravel() for a in ali])
        # dense forward / backward of pair 0 under the final parameters (API parity of forward())
        out['fwd0'] = m.forward(m.vCorpus[0], m.aCorpus[0])
        out['bwd0'] = m.backward(m.vCorpus[0], m.aCorpus[0])
    np.savez_compressed(os.path.join(HERE, 'ik_%s.npz' % name), **out)
    print('wrote ik_%s.npz  avg_ll=%s' % (name, np.array2string(np.array(lls), precision=6)))


def make_ik():
    # (1) the reference's own 3-pair "tiny" sanity corpus (image_phone_hmm_word_discoverer.py:655-667)
    tiny_f = [np.array([[1., 0., 0.], [0., 1., 0.]]), np.array([[0., 1., 0.], [0., 0., 1.]]),
              np.array([[0., 0., 1.], [1., 0., 0.]])]
    tiny_x = [np.array([0, 1]), np.array([1, 2]), np.array([2, 0])]
    run_ik_case('tiny_linear', 'linear', tiny_f, tiny_x, K=3, n_iter=4, seed=1, lr=0.01)
    run_ik_case('tiny_gaussian', 'gaussian', tiny_f, tiny_x, K=3, n_iter=4, seed=2, lr=1.0)

    # (2) short captions, >= 6 distinct n  -> Toeplitz pooling ON, likelihoods above EPS
    rng = np.random.default_rng(20261018)
    f, x = synth_ik_corpus(rng, 24, [1, 2, 3, 4, 5, 6, 7], 2, 14, K=7, P=9, D=6)
    run_ik_case('short_toeplitz_linear', 'linear', f, x, K=7, n_iter=3, seed=3, lr=0.1,
                momentum=0.1, nonuniform_obs=True)
    run_ik_case('short_toeplitz_gaussian', 'gaussian', f, x, K=7, n_iter=3, seed=4, lr=0.1,
                width=2.0, nonuniform_obs=True)

    # (3) MSCOCO-like lengths, < 6 distinct n -> Toeplitz OFF, EPS floors active.  Two iterations
    # only: in this regime the reference's own conceptCounts underflow to 0/0 = NaN at the third.
    rng = np.random.default_rng(20261019)
    f, x = synth_ik_corpus(rng, 16, [3, 5], 30, 46, K=13, P=49, D=16)
    run_ik_case('long_floor_linear', 'linear', f, x, K=13, n_iter=2, seed=5, lr=0.05)
    run_ik_case('long_floor_gaussian', 'gaussian', f, x, K=13, n_iter=2, seed=6, lr=0.05, width=1.0)

    # (4) mixed: some sentences above, some below the EPS likelihood floor; 10 distinct n
    rng = np.random.default_rng(20261020)
    f, x = synth_ik_corpus(rng, 28, list(range(1, 11)), 4, 60, K=9, P=12, D=8)
    run_ik_case('mixed_linear', 'linear', f, x, K=9, n_iter=3, seed=7, lr=0.2, nonuniform_obs=True)
    run_ik_case('mixed_gaussian', 'gaussian', f, x, K=9, n_iter=3, seed=8, lr=0.2, width=4.0)
    # two-layer (ReLU MLP posterior) on the mixed corpus and on the short Toeplitz corpus
    run_ik_case('mixed_twolayer', 'two-layer', f, x, K=9, n_iter=3, seed=9, lr=0.05, hidden_dim=12,
                w_scale=0.4, nonuniform_obs=True)
    rng = np.random.default_rng(20261018)
    f, x = synth_ik_corpus(rng, 24, [1, 2, 3, 4, 5, 6, 7], 2, 14, K=7, P=9, D=6)
    run_ik_case('short_twolayer', 'two-layer', f, x, K=7, n_iter=3, seed=10, lr=0.1, momentum=0.05,
                hidden_dim=150, w_scale=0.3)


if __name__ == '__main__':
    which = sys.argv[1:] or ['ik', 'hmm']
    if 'ik' in which:
        make_ik()
    if 'hmm' in which:
        try:
            from make_golden_hmm import make_hmm
        except ImportError:
            make_hmm = None
        if make_hmm:
            make_hmm()
