"""NumPy restatement of the segment-embedding HMM (config 4): log-domain HMM over
[NULL]+concepts with diagonal-Gaussian(-mixture) emissions on fixed-length segment embeddings.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PINNED PIECE BY PIECE.  hmm/audio_segembed_hmm_word_discoverer.py cannot run end to end in the reference:
it constructs its acoustic model as ``acousticModel(numMixtures, frameDim, fCorpus=..., tCorpus=..., ...)``
(:86-92) but the shipped ``AudioHMMWordDiscoverer`` only accepts a corpus FILE (its matching constructor is
the commented line audio_hmm_word_discoverer.py:15) -> TypeError.  Every numerical piece the class is made
of DOES exist in the reference and is pinned by golden vectors generated from it:
  * segment cutting + resample embedding  : audio_segembed_hmm_word_discoverer.py:52-81,114-155
    -> ``embed`` / ``sent_embeds`` == the reference's unbound ``embed`` / ``getSentEmbeds``
       (tests/golden/seg_pieces.npz, made by tests/golden/make_golden_seg.py)
  * log-domain recursion / counts / Viterbi: audio_hmm_word_discoverer.py:148-254,396-427 with the
    emission hooks of its commented lines :157,:163,:180,:208,:403,:408 (obs_model.logTransProb)
    -> same functions as oracle/plain_hmm.py, pinned by hmm_flickr24_log.npz / hmm_synth_log.npz
  * emission log N(x; mu, diag s2), mixture LSE : smt/audio_gmm_word_discoverer.py:53-106,395-401
    -> ``log_gauss`` / ``emission`` == the reference's ``gaussian(log_prob=True)`` / ``gmmProb(log_prob=True)``
  * mean update as posterior-weighted average   : smt/audio_gmm_word_discoverer.py:339-375
    -> ``weighted_stats`` + ``means_from_stats`` == ``GMMWordDiscoverer.updateTranslationDensities``
What has NO reference counterpart (design choices of the intended model, documented in DESIGN.md): how the
pieces are wired together in ``em_iteration`` (posterior x within-state responsibility as the weight), the
mixture-prior update for numMixtures > 1 and the optional variance update (the reference's GMM class resets
its priors to 0 every M-step, :345-352, and its variance branch reads a stale normaliser, :378-383).
"""
import math

import numpy as np
import scipy.signal as signal
from scipy.special import logsumexp

from . import plain_hmm as ph


def embed(y, embed_dim, frame_dim):
    """embed(), :114-143 (technique='resample'): (frames, featDim) -> (embed_dim,)"""
    y = y[:, :frame_dim].T
    n = int(embed_dim / frame_dim)
    return signal.resample(y.astype('float32'), n, axis=1).flatten('C')


def sent_embeds(x, segmentation, embed_dim, frame_dim):
    """getSentEmbeds(), :145-155"""
    return np.array([embed(x[segmentation[i]:segmentation[i + 1]], embed_dim, frame_dim)
                     for i in range(len(segmentation) - 1)])


def log_gauss(x, mean, var):
    """gaussian(..., cov_type='diag', log_prob=True), smt/audio_gmm_word_discoverer.py:53-61"""
    d = mean.shape[0]
    return -(d / 2. * np.log(2. * math.pi) + np.sum(np.log(var)) / 2.) - np.sum((x - mean) ** 2 / (2. * var), axis=-1)


def emission(x, e, lprior, means, var):
    """lb[t, j] = logTransProb(x_t, e_j) = LSE_m(lprior[w][m] + log N(x_t; mu[w][m], var[w][m]))
    and the within-state mixture responsibilities resp[t, j, m] (log)."""
    T, n, M = x.shape[0], len(e), lprior.shape[1]
    comp = np.zeros((T, n, M))
    for j, w in enumerate(e):
        for m in range(M):
            comp[:, j, m] = lprior[w, m] + log_gauss(x, means[w, m], var[w, m])
    lb = logsumexp(comp, axis=2)
    return lb, comp - lb[:, :, None]


def forward(lb, lpi, lA):
    T, n = lb.shape
    a = -np.inf * np.ones((T, n))
    a[0] = lpi + lb[0]
    for t in range(T - 1):
        for j in range(n):
            a[t + 1, j] = logsumexp(lA[:, j] + a[t]) + lb[t + 1, j]
    return a


def backward(lb, lA):
    T, n = lb.shape
    be = -np.inf * np.ones((T, n))
    be[T - 1] = 0.
    for t in range(T - 1, 0, -1):
        for j in range(n):
            be[t - 1, j] = logsumexp(lA[j] + be[t] + lb[t])
    return be


def estep_pair(x, e, p):
    lb, resp = emission(x, e, p['lprior'], p['means'], p['var'])
    n = len(e)
    lpi, lA = p['init'][n], p['trans'][n]
    a, be = forward(lb, lpi, lA), backward(lb, lA)
    T = lb.shape[0]
    if T < 2:
        raise NameError('transJumpCount')
    init = logsumexp(a + be, axis=0)
    t = T - 2
    E = np.tile(a[t], (n, 1)).T + lA + lb[t + 1] + be[t + 1]
    trans = np.empty((n, n))
    for s in range(n):
        for s2 in range(n):
            trans[s, s2] = logsumexp(np.diagonal(E, offset=s2 - s))
    post = a + be
    post = post - logsumexp(post.flatten())
    return dict(ll=logsumexp(a[-1]), init=init, trans=trans, post=post, resp=resp)


def weighted_stats(embs, tgt, logw, Vt, M):
    """Sufficient statistics of the Gaussian M-step: logw[u] is the (T, n, M) log weight of frame t for
    (state j, mixture m) of utterance u.  Returns (w_sum (Vt, M), x_sum (Vt, M, D), xx_sum (Vt, M, D))."""
    D = embs[0].shape[1]
    w_sum = np.zeros((Vt, M))
    x_sum = np.zeros((Vt, M, D))
    xx_sum = np.zeros((Vt, M, D))
    for x, e, lw in zip(embs, tgt, logw):
        wgt = np.exp(lw)
        for j, w in enumerate(e):
            w_sum[w] += wgt[:, j].sum(0)
            x_sum[w] += wgt[:, j].T @ x
            xx_sum[w] += wgt[:, j].T @ (x ** 2)
    return w_sum, x_sum, xx_sum


def means_from_stats(old_means, w_sum, x_sum):
    """updateTranslationDensities (:339-375): sum_t exp(logw - LSE(logw)) x_t == x_sum / w_sum; a (word,
    mixture) that received no weight keeps its mean."""
    means = old_means.copy()
    with np.errstate(invalid='ignore', divide='ignore'):
        mu = x_sum / w_sum[:, :, None]
    ok = w_sum > 0
    means[ok] = mu[ok]
    return means, mu, ok


def em_iteration(embs, tgt, p, acc, update_var=False):
    """One epoch.  acc: plain_hmm.LogAccumulators-like running init/trans accumulators (the
    reference's count lists live outside the epoch loop); the emission statistics are per epoch."""
    Vt, M, D = p['means'].shape
    lens = sorted(p['init'])
    logw = []
    for x, e in zip(embs, tgt):
        n = len(e)
        r = estep_pair(x, e, p)
        acc.init[n] = np.logaddexp(acc.init[n], r['init'])
        acc.trans[n] = np.logaddexp(acc.trans[n], r['trans'])
        logw.append(r['post'][:, :, None] + r['resp'])           # (T, n, M)
    w_sum, x_sum, xx_sum = weighted_stats(embs, tgt, logw, Vt, M)
    new = dict(p)
    new['init'] = {m: acc.init[m] - logsumexp(acc.init[m]) for m in lens}
    new['trans'] = {m: acc.trans[m] - logsumexp(acc.trans[m], axis=1, keepdims=True) for m in lens}
    seen = w_sum.sum(1) > 0
    means, mu, ok = means_from_stats(p['means'], w_sum, x_sum)
    new['means'] = means
    lprior = p['lprior'].copy()
    if M > 1:
        with np.errstate(divide='ignore'):
            lp = np.log(w_sum / w_sum.sum(1, keepdims=True))
        lprior[seen] = lp[seen]
    new['lprior'] = lprior
    if update_var:
        var = p['var'].copy()
        with np.errstate(invalid='ignore', divide='ignore'):
            v = xx_sum / w_sum[:, :, None] - mu ** 2
        var[ok] = np.maximum(v[ok], 1e-6)
        new['var'] = var
    ll = 0.
    for x, e in zip(embs, tgt):
        n = len(e)
        lb, _ = emission(x, e, new['lprior'], new['means'], new['var'])
        ll += logsumexp(forward(lb, new['init'][n], new['trans'][n])[-1])
    return new, dict(avg_ll=ll / len(embs), w_sum=w_sum)


def align(x, e, p):
    """align with the emission hooks of audio_hmm_word_discoverer.py:396-427"""
    lb, _ = emission(x, e, p['lprior'], p['means'], p['var'])
    n, T = len(e), lb.shape[0]
    lpi, lA = p['init'][n], p['trans'][n]
    scores = lpi + lb[0]
    bp = np.zeros((T, n), dtype=int)
    probs = []
    for t in range(1, T):
        cand = np.tile(scores, (n, 1)).T + lA + lb[t]
        bp[t] = np.argmax(cand, axis=0)
        scores = np.max(cand, axis=0)
        probs.append(scores.tolist())
    cur = int(np.argmax(scores))
    path = [cur]
    for t in range(T - 1, 0, -1):
        cur = int(bp[t, cur])
        path.append(cur)
    return path[::-1], probs
