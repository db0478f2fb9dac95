"""NumPy float64 restatement of the (region i, concept k)-state HMM word discoverers.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity is pinned against the unmodified
reference classes via ``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``.

Reference (all paths relative to /root/reference):
  linear   : hmm_dnn/image_phone_hmm_word_discoverer.py            (``ImagePhoneHMMWordDiscoverer``)
  gaussian : hmm_dnn/image_phone_gaussian_hmm_word_discoverer.py   (``ImagePhoneGaussianHMMWordDiscoverer``)

A corpus here is ``(feats, phones)``: ``feats[ex]`` is an (n, D) float64 array of region features,
``phones[ex]`` an int array of length T of phone ids (the reference's one-hot ``aSen`` rows collapse
to their index: ``obs @ aSen[t] == obs[:, x_t]`` exactly).
"""
import math

import numpy as np
from scipy.special import logsumexp

EPS = 1e-50  # hmm_dnn/image_phone_hmm_word_discoverer.py:11


# ----------------------------------------------------------------------------------------------
# image posterior  p(z_i = k | v_i)
# ----------------------------------------------------------------------------------------------
def posterior_linear(v, W):
    """softmaxLayer, image_phone_hmm_word_discoverer.py:533-541."""
    vc = np.concatenate([v, np.ones((v.shape[0], 1))], axis=1)
    x = vc @ W.T
    return np.exp(x.T - logsumexp(x, axis=1)).T


def posterior_gaussian(v, mus, width):
    """softmaxLayer, image_phone_gaussian_hmm_word_discoverer.py:501-510."""
    x = np.zeros((v.shape[0], mus.shape[0]))
    for i in range(v.shape[0]):
        x[i] = -np.sum((v[i] - mus) ** 2, axis=1) / width
    return np.exp(x.T - logsumexp(x, axis=1)).T


# ----------------------------------------------------------------------------------------------
# forward / backward  (T, n, K) in the raw probability domain, no rescaling
# ----------------------------------------------------------------------------------------------
def forward(pz, x, obs, pi, A):
    """forward, image_phone_hmm_word_discoverer.py:276-304."""
    T, (n, K) = len(x), pz.shape
    d = np.diag(A)
    Aoff = A - np.diag(d)
    a = np.zeros((T, n, K))
    a[0] = pi[:, None] * pz * obs[:, x[0]]
    for t in range(T - 1):
        o = obs[:, x[t + 1]]
        a[t + 1] = (d[:, None] * a[t]) * o
        a[t + 1] += ((Aoff.T @ a[t].sum(-1)) * (pz * o).T).T
    return a


def backward(pz, x, obs, A):
    """backward, image_phone_hmm_word_discoverer.py:314-335."""
    T, (n, K) = len(x), pz.shape
    d = np.diag(A)
    Aoff = A - np.diag(d)
    b = np.zeros((T, n, K))
    b[T - 1] = 1.0
    for t in range(T - 1, 0, -1):
        o = obs[:, x[t]]
        b[t - 1] = d[:, None] * (b[t] * o)
        b[t - 1] += (Aoff @ np.sum(b[t] * (pz * o), axis=-1))[:, None]
    return b


def pair_loglik(a):
    """computeAvgLogLikelihood body, :523-531 (floored at EPS)."""
    return math.log(max(float(np.sum(a[-1])), EPS))


# ----------------------------------------------------------------------------------------------
# expected counts
# ----------------------------------------------------------------------------------------------
def init_counts(a, b):
    """updateInitialCounts, :347-362 -- occupancy summed over ALL t, elementwise EPS floor."""
    g = np.maximum(a * b, EPS)                       # (T, n, K)
    return np.sum(g.sum(-1) / g.sum((1, 2))[:, None], axis=0)


def trans_counts(a, b, pz, x, obs, A, toeplitz):
    """updateTransitionCounts, :374-416."""
    T, n = a.shape[0], a.shape[1]
    d = np.diag(A)
    Aoff = A - np.diag(d)
    out = np.zeros((n, n))
    for t in range(T - 1):
        o = obs[:, x[t + 1]]
        xi = np.diag(np.sum(a[t] * d[:, None] * o * b[t + 1], axis=-1))
        xi = xi + a[t].sum(-1)[:, None] * Aoff * np.sum(pz * o * b[t + 1], axis=-1)
        xi = np.maximum(xi, EPS)
        xi = xi / xi.sum()
        if toeplitz:
            xi = toeplitz_pool(xi)
        out += xi
    return out


def toeplitz_pool(xi):
    """:399-413 -- every [s][s'] receives the sum of its diagonal s'-s."""
    n = xi.shape[0]
    out = np.empty_like(xi)
    for s in range(n):
        for s2 in range(n):
            out[s, s2] = np.trace(xi, offset=s2 - s)
    return out


def state_counts(a, b):
    """updateStateCounts, :426-433 -- the NORMALISER is floored, so gamma is un-normalised
    (= alpha*beta/EPS) whenever the sentence likelihood is below EPS."""
    g = a * b
    return g / np.maximum(g.sum((1, 2)), EPS)[:, None, None]


def concept_counts(pz, x, obs, pi, A):
    """updateConceptCounts, :443-465.  For each (i,k): clamp region i to concept k and run a
    plain n-state forward with marginal emissions; no floor anywhere."""
    n, K = pz.shape
    T = len(x)
    o = obs[:, x]                                    # (K, T)
    e = pz @ o                                       # (n, T) marginal emission
    # F[i, k, j]; emission e'[i,k,j,t] = e[j,t] for j != i, o[k,t] for j == i
    ii = np.arange(n)
    F = np.broadcast_to(pi * e[:, 0], (n, K, n)).copy()
    F[ii, :, ii] = pi[:, None] * o[:, 0][None, :]
    for t in range(1, T):
        F = F @ A
        em = np.broadcast_to(e[:, t], (n, K, n)).copy()
        em[ii, :, ii] = o[:, t][None, :]
        F = F * em
    lik = F.sum(-1)                                  # (n, K)
    num = pz * lik
    return (num.T / np.sum(num, axis=1)).T


# ----------------------------------------------------------------------------------------------
# one EM iteration over a corpus
# ----------------------------------------------------------------------------------------------
def estep_pair(v, x, params, kind):
    """Everything trainUsingEM computes for one pair (:224-235). Returns a dict."""
    n = v.shape[0]
    pz = posterior(v, params, kind)
    pi, A, obs = params['init'][n], params['trans'][n], params['obs']
    a = forward(pz, x, obs, pi, A)
    b = backward(pz, x, obs, A)
    gam = state_counts(a, b)
    cA = gam.sum(1)                                  # (T, K)  conceptCountsA[ex]
    return dict(
        pz=pz,
        ll=pair_loglik(a),
        init=init_counts(a, b),
        trans=trans_counts(a, b, pz, x, obs, A, params['toeplitz']),
        cA=cA,
        cC=concept_counts(pz, x, obs, pi, A),
    )


def hidden_layer(v, V):
    """hiddenLayer (ReLU), image_phone_hmm_dnn_word_discoverer.py:573-579."""
    vc = np.concatenate([v, np.ones((v.shape[0], 1))], axis=1)
    h = vc @ V.T
    return h * (h > 0.)


def posterior(v, params, kind):
    if kind == 'linear':
        return posterior_linear(v, params['W'])
    if kind == 'two-layer':   # softmaxLayer(hiddenLayer(v)), image_phone_hmm_dnn_word_discoverer.py:310,581-590
        return posterior_linear(hidden_layer(v, params['V']), params['W'])
    return posterior_gaussian(v, params['mus'], params['width'])


def em_iteration(feats, phones, params, kind='linear', update_lr=False, epoch=0):
    """One epoch body of trainUsingEM (:207-261; gaussian :204-264).

    ``params``: dict(init={m:(m,)}, trans={m:(m,m)}, obs=(K,P), W=(K,D+1) | mus=(K,D), width,
    lr, momentum, toeplitz=bool).  Returns (new_params, info) where info carries the average
    log-likelihood of the *entering* parameters, the raw count tables and the per-pair outputs.
    """
    K, P = params['obs'].shape
    N = len(feats)
    lens = sorted(params['init'].keys())
    initC = {m: np.zeros((m,)) for m in lens}
    transC = {m: np.zeros((m, m)) for m in lens}
    phoneC = np.zeros((K, P))
    cC_all, cA_all, pz_all = [], [], []
    ll = 0.0
    for v, x in zip(feats, phones):
        r = estep_pair(v, x, params, kind)
        n = v.shape[0]
        ll += r['ll']
        initC[n] += r['init']
        transC[n] += r['trans']
        # phoneCounts += sum_i(gamma).T @ onehot  (:233)
        np.add.at(phoneC.T, x, r['cA'])
        cC_all.append(r['cC'])
        cA_all.append(r['cA'])
        pz_all.append(r['pz'])

    new = dict(params)
    new['init'], new['trans'] = {}, {}
    for m in lens:
        if kind == 'two-layer' and params.get('freeze_trans', False):
            # image_phone_hmm_dnn_word_discoverer.py:247-273 with freezeTransition=True
            ic = np.maximum(initC[m], EPS)
            new['init'][m] = ic / np.sum(ic)
            new['trans'][m] = params['trans'][m].copy()
            continue
        if kind == 'linear':
            new['init'][m] = initC[m] / np.sum(initC[m])                      # :240
            tot = np.sum(transC[m], axis=1)                                   # :245
            tr = params['trans'][m].copy()
            for s in range(m):
                if tot[s] != 0:
                    tr[s] = transC[m][s] / tot[s]
            new['trans'][m] = tr
        else:
            ic = np.maximum(initC[m], EPS)                                    # gaussian :239
            new['init'][m] = ic / np.sum(ic)
            tot = np.sum(np.maximum(transC[m], EPS), axis=1)                  # gaussian :242
            tr = params['trans'][m].copy()
            for s in range(m):
                if tot[s] != 0:
                    tr[s] = np.maximum(transC[m][s], EPS) / tot[s]
            new['trans'][m] = tr
    norm = np.sum(np.maximum(phoneC, EPS), axis=-1)                           # :255
    new['obs'] = (phoneC.T / norm).T

    lr, mom = params['lr'], params['momentum']
    if kind == 'linear':
        # updateSoftmaxWeight :475-488 (pz recomputed with the OLD W == pz_all)
        D = feats[0].shape[1]
        dW = np.zeros((K, D + 1))
        for v, cC, pz in zip(feats, cC_all, pz_all):
            vc = np.concatenate([v, np.ones((v.shape[0], 1))], axis=1)
            dW += 1.0 / N * (cC - pz).T @ vc
        new['W'] = (1.0 - mom) * params['W'] + lr * dW
        grad = dW
    elif kind == 'two-layer':
        # updateNeuralNetWeights, image_phone_hmm_dnn_word_discoverer.py:504-528
        Wm, Vm = params['W'], params['V']
        dW = np.zeros_like(Wm)
        dV = np.zeros_like(Vm)
        for v, cC, pz in zip(feats, cC_all, pz_all):
            h = hidden_layer(v, Vm)
            Delta = cC - pz
            Eps = Delta @ Wm[:, :-1]
            vc = np.concatenate([v, np.ones((v.shape[0], 1))], axis=1)
            hc = np.concatenate([h, np.ones((h.shape[0], 1))], axis=1)
            dW += 1.0 / N * Delta.T @ hc
            dV += 1.0 / N * (Eps * (h > 0)).T @ vc
        new['W'] = (1.0 - mom) * Wm + lr * dW
        new['V'] = (1.0 - mom) * Vm + lr * dV
        grad = dW
    else:
        # gaussian updateSoftmaxWeight :488-499 (non-exact branch)
        mus, width = params['mus'], params['width']
        dmus = np.zeros_like(mus)
        for v, cC, pz in zip(feats, cC_all, pz_all):
            Delta = cC - pz
            dmus += 1.0 / (N * width) * (Delta.T @ v - (np.sum(Delta, axis=0) * mus.T).T)
        new['mus'] = (1.0 - mom) * mus + lr * dmus
        grad = dmus
    if update_lr and (epoch + 1) % 10 == 0:                                   # :260-261
        new['lr'] = lr / 10
    info = dict(avg_ll=ll / N, initC=initC, transC=transC, phoneC=phoneC, cC=cC_all, cA=cA_all,
                pz=pz_all, grad=grad)
    return new, info


def initial_params(feats, n_words, n_phones, kind='linear', W=None, mus=None, width=1.0, lr=10.0,
                   momentum=0.0, obs=None):
    """initializeModel, :106-147 (uniform init/trans/obs; W or mus must be injected --
    the reference draws W from the global RNG / KMeans, which the oracle does not imitate)."""
    lens = sorted({v.shape[0] for v in feats})
    p = dict(
        init={m: np.ones((m,)) / m for m in lens},
        trans={m: np.ones((m, m)) / m for m in lens},
        obs=(np.ones((n_words, n_phones)) / n_phones) if obs is None else np.array(obs, dtype=float),
        lr=lr, momentum=momentum,
        toeplitz=len(lens) >= 6,                                              # :399
    )
    if kind == 'linear':
        p['W'] = np.array(W, dtype=float)
    elif kind == 'two-layer':
        p['W'] = np.array(W, dtype=float)
        p['V'] = np.array(mus, dtype=float)        # hidden weights travel in the `mus` slot
    else:
        p['mus'] = np.array(mus, dtype=float)
        p['width'] = width
    return p


# ----------------------------------------------------------------------------------------------
# decoding
# ----------------------------------------------------------------------------------------------
def align(pz, x, obs, pi, A, floor_norm=False, floor_scores=True):
    """align, :543-584 (Viterbi over regions with marginal emissions, EPS score floor).
    ``floor_norm`` selects the gaussian class's floored alignProbs normaliser (gaussian :583);
    ``floor_scores=False`` is the two-layer class, which does not floor the scores
    (image_phone_hmm_dnn_word_discoverer.py:612)."""
    n = pz.shape[0]
    T = len(x)
    onehot = np.zeros((T, obs.shape[1]))
    onehot[np.arange(T), x] = 1.0
    p = (pz @ (obs @ onehot.T)).T                    # (T, n) -- same BLAS call shape as :550
    bp = np.zeros((T, n), dtype=int)
    scores = pi * p[0]
    probs = [scores.tolist()]
    for t in range(1, T):
        cand = np.tile(scores, (n, 1)).T * A * p[t]
        bp[t] = np.argmax(cand, axis=0)
        scores = np.max(cand, axis=0)
        if floor_scores:
            scores = np.maximum(scores, EPS)
        if floor_norm:
            probs.append((scores / np.sum(np.maximum(scores, EPS))).tolist())
        else:
            probs.append((scores / np.sum(scores)).tolist())
    cur = int(np.argmax(scores))
    path = [cur]
    for t in range(T - 1, 0, -1):
        cur = int(bp[t, cur])
        path.append(cur)
    return path[::-1], probs


def cluster(pz, x, obs, alignment):
    """cluster, :586-597."""
    n = pz.shape[0]
    scores = np.array(pz, dtype=float, copy=True)
    for i in range(n):
        for t in range(len(x)):
            if alignment[t] == i:
                scores[i] *= obs[:, x[t]]
    return np.argmax(scores, axis=1).tolist(), scores
