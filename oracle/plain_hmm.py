"""NumPy float64 restatement of the plain-state HMM word discoverers of hmm/.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned against the unmodified reference
classes via ``tests/golden/make_golden_hmm.py`` -> ``tests/golden/hmm_*.npz``.

Reference (paths relative to /root/reference):
  prob : hmm/hmm_word_discoverer.py        (``HMMWordDiscoverer``, probability domain, no floors)
  log  : hmm/audio_hmm_word_discoverer.py  (``AudioHMMWordDiscoverer``, log domain, NULL state;
         hmm/hmm_word_discoverer_logscale.py is the same file with un-importable imports)

States are the concept tokens of the caption (the log class prepends NULL).  The dict-of-dict
``obs[tw][fw]`` becomes a dense (Vt, Vf) table with NaN marking pairs that never co-occur
("not in the dict"); sentences are integer id arrays.
"""
import math

import numpy as np
from scipy.special import logsumexp

UNK = 10e-12  # align()'s unkProb default (hmm_word_discoverer.py:301)


# ----------------------------------------------------------------------------------------------
# shared helpers
# ----------------------------------------------------------------------------------------------
def cooccurrence_mask(tgt, src, Vt, Vf):
    seen = np.zeros((Vt, Vf), dtype=bool)
    for e, f in zip(tgt, src):
        seen[np.ix_(np.unique(e), np.unique(f))] = True
    return seen


def toeplitz_pool(xi):
    n = xi.shape[0]
    out = np.empty_like(xi)
    for s in range(n):
        for s2 in range(n):
            out[s, s2] = np.trace(xi, offset=s2 - s)
    return out


# ----------------------------------------------------------------------------------------------
# probability-domain class
# ----------------------------------------------------------------------------------------------
def prob_initial_obs(tgt, src, Vt, Vf):
    """initializeModel, hmm_word_discoverer.py:90-108: co-occurrence counts (every (tw, fw) token
    pair of every sentence pair counts once), row-normalised; absent pairs are NaN."""
    c = np.zeros((Vt, Vf))
    for e, f in zip(tgt, src):
        fe = np.bincount(e, minlength=Vt).astype(float)
        ff = np.bincount(f, minlength=Vf).astype(float)
        c += np.outer(fe, ff)
    obs = np.full((Vt, Vf), np.nan)
    rows = c.sum(1) > 0
    with np.errstate(invalid='ignore', divide='ignore'):
        norm = c / c.sum(1, keepdims=True)
    obs[c > 0] = norm[c > 0]
    return obs


def prob_emis(obs, e, f, unseen=0.0):
    """b[t, j] = obs[e_j][f_t] if present else `unseen` (:122)."""
    b = obs[np.ix_(e, f)].T.copy()
    b[np.isnan(b)] = unseen
    return b


def prob_forward(e, f, obs, pi, A):
    """forward, :110-125"""
    b = prob_emis(obs, e, f)
    T, n = b.shape
    a = np.zeros((T, n))
    a[0] = pi * b[0]
    for t in range(T - 1):
        a[t + 1] = A.T @ a[t] * b[t + 1]
    return a


def prob_backward(e, f, obs, A):
    """backward, :127-138"""
    b = prob_emis(obs, e, f)
    T, n = b.shape
    be = np.zeros((T, n))
    be[T - 1] = 1.0
    for t in range(T - 1, 0, -1):
        be[t - 1] = A @ (be[t] * b[t])
    return be


def prob_estep_pair(e, f, obs, pi, A):
    b = prob_emis(obs, e, f)
    a = prob_forward(e, f, obs, pi, A)
    be = prob_backward(e, f, obs, A)
    T, n = a.shape
    g = a * be
    gam = g / g.sum(1, keepdims=True)                          # :147-148, :193
    trans = np.zeros((n, n))
    for t in range(T - 1):
        xi = np.tile(a[t], (n, 1)).T * (b[t + 1] * be[t + 1]) * A   # :161
        xi = xi / xi.sum()
        trans += toeplitz_pool(xi)                              # :165-177 (always pooled)
    return dict(ll=math.log(a[-1].sum()), init=gam.sum(0), trans=trans, gam=gam)


def prob_em_iteration(tgt, src, params):
    """One epoch body of trainUsingEM (:254-296).  params: init{m}, trans{m}, obs (Vt,Vf).
    Returns (new params, info) -- info['avg_ll'] is the LL of the ENTERING parameters (:259-260)."""
    obs = params['obs']
    lens = sorted(params['init'])
    initC = {m: np.zeros(m) for m in lens}
    transC = {m: np.zeros((m, m)) for m in lens}
    obsC = np.zeros_like(obs)
    ll = 0.0
    for e, f in zip(tgt, src):
        n = len(e)
        r = prob_estep_pair(e, f, obs, params['init'][n], params['trans'][n])
        ll += r['ll']
        initC[n] += r['init']
        transC[n] += r['trans']
        for t in range(len(f)):
            for i in range(n):
                obsC[e[i], f[t]] += r['gam'][t, i]
    new = dict(init={}, trans={}, obs=None)
    for m in lens:
        new['init'][m] = initC[m] / initC[m].sum()
        tot = transC[m].sum(1)
        tr = params['trans'][m].copy()
        for s in range(m):
            if tot[s] != 0:
                tr[s] = transC[m][s] / tot[s]
        new['trans'][m] = tr
    present = ~np.isnan(obs)
    norm = np.where(present, obsC, 0.0).sum(1, keepdims=True)
    with np.errstate(invalid='ignore', divide='ignore'):
        o = obsC / norm
    new['obs'] = np.where(present, o, np.nan)
    return new, dict(avg_ll=ll / len(tgt), initC=initC, transC=transC, obsC=obsC)


def prob_align(e, f, obs, pi, A, unk=UNK):
    """align, :301-329 (no floor; alignProbs start at t=1)."""
    n, T = len(e), len(f)
    scores = pi * obs[e, f[0]]
    bp = np.zeros((T, n), dtype=int)
    probs = []
    for t in range(1, T):
        b = obs[e, f[t]].copy()
        b[np.isnan(b)] = unk
        cand = np.tile(scores, (n, 1)).T * A * b
        bp[t] = np.argmax(cand, axis=0)
        scores = np.max(cand, axis=0)
        probs.append((scores / np.sum(scores)).tolist())
    cur = int(np.argmax(scores))
    path = [cur]
    for t in range(T - 1, 0, -1):
        cur = int(bp[t, cur])
        path.append(cur)
    return path[::-1], probs


# ----------------------------------------------------------------------------------------------
# log-domain class (targets already carry the NULL state in position 0)
# ----------------------------------------------------------------------------------------------
def log_initial_obs(tgt, src, Vt, Vf):
    """initializeModel, audio_hmm_word_discoverer.py:123-138: every co-occurring (tw, fw) gets
    count 1 (not accumulated), row-normalised, log."""
    seen = cooccurrence_mask(tgt, src, Vt, Vf)
    obs = np.full((Vt, Vf), np.nan)
    cnt = seen.sum(1, keepdims=True).astype(float)
    with np.errstate(divide='ignore', invalid='ignore'):
        val = np.log(1.0 / cnt) * np.ones((1, Vf))
    obs[seen] = val[seen]
    return obs


def log_emis(obs, e, f, unseen=0.0):
    b = obs[np.ix_(e, f)].T.copy()
    b[np.isnan(b)] = unseen                                     # ":161 ... else 0"
    return b


def log_forward(e, f, obs, lpi, lA):
    """forward, :148-168"""
    b = log_emis(obs, e, f)
    T, n = b.shape
    a = -np.inf * np.ones((T, n))
    a[0] = lpi + b[0]
    for t in range(T - 1):
        for j in range(n):
            a[t + 1, j] = logsumexp(lA[:, j] + a[t]) + b[t + 1, j]
    return a


def log_backward(e, f, obs, lA):
    """backward, :170-185"""
    b = log_emis(obs, e, f)
    T, n = b.shape
    be = -np.inf * np.ones((T, n))
    be[T - 1] = 0.0
    for t in range(T - 1, 0, -1):
        for j in range(n):
            be[t - 1, j] = logsumexp(lA[j] + be[t] + b[t])
    return be


def log_estep_pair(e, f, obs, lpi, lA):
    """Per-pair log counts (:187-254).  Transition counts use ONLY the last t (the reference's
    transJumpCount dict is re-created inside the t loop, :212) and are Toeplitz-pooled by LSE."""
    b = log_emis(obs, e, f)
    a = log_forward(e, f, obs, lpi, lA)
    be = log_backward(e, f, obs, lA)
    T, n = a.shape
    if T < 2:
        raise NameError('transJumpCount')                       # the reference fails the same way
    init = logsumexp(a + be, axis=0)                            # :192-194 (un-normalised)
    t = T - 2
    E = np.tile(a[t], (n, 1)).T + lA + b[t + 1] + be[t + 1]     # :209
    trans = np.empty((n, n))
    for s in range(n):
        for s2 in range(n):
            trans[s, s2] = logsumexp(np.diagonal(E, offset=s2 - s))
    post = a + be
    post = post - logsumexp(post.flatten())                     # :244-246 (global normaliser)
    return dict(ll=logsumexp(a[-1]), init=init, trans=trans, post=post)


class LogAccumulators(object):
    """The reference keeps every per-pair count in Python lists created OUTSIDE the epoch loop
    (:322-325), so counts accumulate over epochs.  LSE is associative, so running
    log-accumulators are equivalent."""

    def __init__(self, lens, Vt, Vf):
        self.init = {m: np.full(m, -np.inf) for m in lens}
        self.trans = {m: np.full((m, m), -np.inf) for m in lens}
        self.obs = np.full((Vt, Vf), -np.inf)


def log_em_iteration(tgt, src, params, acc):
    """One epoch body of trainUsingEM (:327-391).  Returns (new params, info); info['avg_ll'] is the
    LL of the UPDATED parameters (the reference prints it after the M-step, :391)."""
    obs = params['obs']
    lens = sorted(params['init'])
    for e, f in zip(tgt, src):
        n = len(e)
        r = log_estep_pair(e, f, obs, params['init'][n], params['trans'][n])
        acc.init[n] = np.logaddexp(acc.init[n], r['init'])
        acc.trans[n] = np.logaddexp(acc.trans[n], r['trans'])
        # :248-252,:340 -- per pair, LSE over the occurrences of each (tw, fw); then appended
        pair = {}
        for t in range(len(f)):
            for i in range(n):
                key = (e[i], f[t])
                pair[key] = np.logaddexp(pair.get(key, -np.inf), r['post'][t, i])
        for (tw, fw), v in pair.items():
            acc.obs[tw, fw] = np.logaddexp(acc.obs[tw, fw], v)
    new = dict(init={}, trans={}, obs=None)
    for m in lens:
        new['init'][m] = acc.init[m] - logsumexp(acc.init[m])                 # :355-360
        new['trans'][m] = acc.trans[m] - logsumexp(acc.trans[m], axis=1, keepdims=True)  # :364-369
    present = ~np.isnan(obs)
    masked = np.where(present, acc.obs, -np.inf)
    with np.errstate(invalid='ignore'):
        o = acc.obs - logsumexp(masked, axis=1, keepdims=True)                # :373-389
    new['obs'] = np.where(present, o, np.nan)
    ll = 0.0
    for e, f in zip(tgt, src):
        n = len(e)
        ll += logsumexp(log_forward(e, f, new['obs'], new['init'][n], new['trans'][n])[-1])
    return new, dict(avg_ll=ll / len(tgt))


def log_align(e, f, obs, lpi, lA, unk=UNK):
    """align, :396-427 (alignProbs are the raw log scores from t=1)."""
    n, T = len(e), len(f)
    scores = lpi + obs[e, f[0]]
    bp = np.zeros((T, n), dtype=int)
    probs = []
    for t in range(1, T):
        b = obs[e, f[t]].copy()
        b[np.isnan(b)] = unk
        cand = np.tile(scores, (n, 1)).T + lA + b
        bp[t] = np.argmax(cand, axis=0)
        scores = np.max(cand, axis=0)
        probs.append(scores.tolist())
    cur = int(np.argmax(scores))
    path = [cur]
    for t in range(T - 1, 0, -1):
        cur = int(bp[t, cur])
        path.append(cur)
    return path[::-1], probs
