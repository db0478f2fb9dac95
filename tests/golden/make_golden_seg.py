#!/usr/bin/env python
"""Golden vectors that pin oracle/segembed_hmm.py (config 4) to code the reference DOES ship.

    python tests/golden/make_golden_seg.py        # build container only (needs /root/reference)

``SegEmbedHMMWordDiscoverer`` cannot run end to end in the reference (its acoustic-model constructor call
raises TypeError), but every numerical piece it is made of exists and is importable:

  * ``gaussian(..., log_prob=True)`` and ``gmmProb(..., log_prob=True)``
    (smt/audio_gmm_word_discoverer.py:53-106; the module imports once ``nltk`` is stubbed -- shim_stubs/);
  * the posterior-weighted mean update ``GMMWordDiscoverer.updateTranslationDensities`` (:339-375), driven
    here on an object created without its file-reading constructor, from given log-posteriors;
  * ``SegEmbedHMMWordDiscoverer.embed / getSentEmbeds`` (hmm/audio_segembed_hmm_word_discoverer.py:114-155),
    called unbound on an object created without its constructor (the constructor is what is broken).

The log-domain recursion / counts / Viterbi the class reuses are pinned separately by hmm_*_log.npz.
Writes tests/golden/seg_pieces.npz: inputs + the reference's outputs.
"""
import contextlib
import importlib.util
import io
import os
import sys

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def load_ref(relpath, alias):
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


def main():
    sys.path.insert(0, os.path.join(ROOT, 'shim_stubs'))     # nltk stub (absent from the image)
    sys.path.insert(0, REF)
    gm = load_ref('smt/audio_gmm_word_discoverer.py', 'ref_audio_gmm')
    se = load_ref('hmm/audio_segembed_hmm_word_discoverer.py', 'ref_segembed')
    rng = np.random.default_rng(20261102)
    out = {}

    # ---- (1) gaussian(log_prob=True) and gmmProb(log_prob=True) on (S x D) blocks
    D, M, S = 120, 3, 17
    x = (0.5 * rng.standard_normal((S, D))).astype(np.float32).astype(np.float64)
    means = 0.5 * rng.standard_normal((M, D))
    var = 0.02 + 0.3 * rng.random((M, D))
    lprior = np.log(rng.dirichlet(np.ones(M)))
    out['g_x'], out['g_means'], out['g_var'], out['g_lprior'] = x, means, var, lprior
    out['g_loggauss'] = np.stack([gm.gaussian(x, means[m], np.diag(var[m]), log_prob=True) for m in range(M)])
    out['g_gmm'] = gm.gmmProb(x, lprior, means, var, log_prob=True)
    out['g_gmm_row0'] = gm.gmmProb(x[0], lprior, means, var, log_prob=True)
    # fixed variance 0.02 (the wrapper's fixedVariance default, audio_segembed_hmm_word_discoverer.py:52)
    var_fix = 0.02 * np.ones((M, D))
    out['g_gmm_fixedvar'] = gm.gmmProb(x, lprior, means, var_fix, log_prob=True)

    # ---- (2) posterior-weighted mean update (:339-375) from given log-posteriors
    words = ['NULL', 'dog', 'ball', 'tree']
    Mw = 2
    Dm = 24
    tC, fC, ali = [], [], []
    for u in range(7):
        n = int(rng.integers(1, 4))
        ts = ['NULL'] + [words[int(k)] for k in rng.integers(1, len(words), n)]     # repeats allowed
        T = int(rng.integers(2, 9))
        fs = rng.standard_normal((T, Dm))
        a = {}
        for k_t, tw in enumerate(ts):
            a[str(k_t) + '_' + tw] = np.log(rng.random((Mw, T)) + 1e-3) - 2.0
        tC.append(ts)
        fC.append(fs)
        ali.append(a)
    g = object.__new__(gm.GMMWordDiscoverer)
    g.tCorpus, g.fCorpus, g.alignProb = tC, fC, ali
    g.numMixtures = {w: Mw for w in words}
    g.featDim = Dm
    g.fixedVariance = 0.02
    with contextlib.redirect_stdout(io.StringIO()):
        g.updateTranslationDensities()
    out['m_words'] = np.array(words)
    out['m_n_utts'] = len(tC)
    out['m_tgt'] = np.concatenate([[words.index(w) for w in ts] for ts in tC])
    out['m_tgt_off'] = np.cumsum([0] + [len(ts) for ts in tC])
    out['m_x'] = np.concatenate(fC)
    out['m_x_off'] = np.cumsum([0] + [len(fs) for fs in fC])
    # log weights in (utt, state, mixture, t) order, flattened
    out['m_logw'] = np.concatenate([ali[u][str(k) + '_' + tw].ravel() for u, ts in enumerate(tC) for k, tw in enumerate(ts)])
    out['m_means'] = np.stack([g.transMeans[w] for w in words])

    # ---- (3) embed / getSentEmbeds (:114-155)
    obj = object.__new__(se.SegEmbedHMMWordDiscoverer)
    obj.embedDim, obj.frameDim, obj.featDim = 120, 12, 14
    utt = rng.standard_normal((83, 14))
    seg = [0, 7, 19, 20, 41, 60, 83]               # includes a 1-frame segment
    out['e_utt'], out['e_seg'] = utt, np.array(seg)
    out['e_embeds'] = obj.getSentEmbeds(utt, seg, frameDim=12)
    out['e_embed_first'] = obj.embed(utt[0:7], frameDim=12)
    obj2 = object.__new__(se.SegEmbedHMMWordDiscoverer)
    obj2.embedDim, obj2.frameDim, obj2.featDim = 560, 12, 14     # run_audio.py:136 (mscoco2k)
    out['e_embeds_560'] = obj2.getSentEmbeds(utt, seg, frameDim=12)
    np.savez_compressed(os.path.join(HERE, 'seg_pieces.npz'), **out)
    print('wrote seg_pieces.npz:', {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == '__main__':
    main()
