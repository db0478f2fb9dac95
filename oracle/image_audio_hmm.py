"""NumPy float64 restatement of ``ImageAudioHMMWordDiscoverer``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity is pinned against the unmodified
reference class via ``tests/golden/make_golden_audio.py`` -> ``tests/golden/ia_*.npz``.

Reference: hmm_dnn/image_audio_hmm_word_discoverer.py (paths relative to /root/reference).  The class
is the image-phone HMM (``oracle/image_phone_hmm.py``) with the discrete phone emission
``obs[:, x_t]`` replaced by a DENSE per-frame emission

    E[t, k] = sum_ph phoneProbs[k, ph] * softmax_ph(WA . [a_t; 1])          (:286-288)

so every recursion below is the image-phone one evaluated with ``obs = E.T`` and ``x = arange(T)``.
What is new: ``updateConceptPhoneCounts`` (:486-493), the phone-probability M-step (:250-251) and
the audio posterior update (:527-541).
"""
import math

import numpy as np

from . import image_phone_hmm as ip

EPS = ip.EPS  # hmm_dnn/image_audio_hmm_word_discoverer.py:11


def emissions(a, WA, phone_probs):
    """softmaxLayerA (:549-554) and probs_x_given_z (:288).  Returns (ph (T, nPhones), E (T, K))."""
    ph = ip.posterior_linear(a, WA)
    return ph, (phone_probs @ ph.T).T


def concept_phone_counts(fwd, bwd, ph):
    """updateConceptPhoneCounts, :486-493: per t the outer product of sum_i alpha_t beta_t (K,) and
    p(ph | x_t) (nPhones,), normalised to sum 1 (no floor: NaN if the posterior mass is 0)."""
    T = ph.shape[0]
    out = np.zeros((T, fwd.shape[2], ph.shape[1]))
    for t in range(T):
        out[t] = np.sum(fwd[t, np.newaxis] * bwd[t, np.newaxis], axis=1).T @ ph[t, np.newaxis]
        out[t] /= np.sum(out[t])
    return out


def estep_pair(v, a, params):
    """Everything trainUsingEM computes for one pair (:221-232)."""
    n = v.shape[0]
    T = a.shape[0]
    pz = ip.posterior_linear(v, params['WV'])
    ph, E = emissions(a, params['WA'], params['phone_probs'])
    x = np.arange(T)
    obs = np.ascontiguousarray(E.T)                     # obs[:, x_t] == E[t]
    pi, A = params['init'][n], params['trans'][n]
    fwd = ip.forward(pz, x, obs, pi, A)
    bwd = ip.backward(pz, x, obs, A)
    return dict(
        pz=pz, ph=ph, E=E,
        ll=ip.pair_loglik(fwd),
        init=ip.init_counts(fwd, bwd),
        trans=ip.trans_counts(fwd, bwd, pz, x, obs, A, params['toeplitz']),
        cpc=concept_phone_counts(fwd, bwd, ph),
        cC=ip.concept_counts(pz, x, obs, pi, A),
    )


def em_iteration(feats, audio, params, update_lr=False, epoch=0):
    """One epoch body of trainUsingEM (:203-262).  ``params``: init, trans, phone_probs (K, nPhones),
    WV (K, D+1), WA (nPhones, Da+1), lr, momentum, toeplitz."""
    K, nPh = params['phone_probs'].shape
    N = len(feats)
    lens = sorted(params['init'].keys())
    initC = {m: np.zeros((m,)) for m in lens}
    transC = {m: np.zeros((m, m)) for m in lens}
    phoneC = np.zeros((K, nPh))
    cC_all, pz_all, ph_all, cpc_all = [], [], [], []
    ll = 0.0
    for v, a in zip(feats, audio):
        r = estep_pair(v, a, params)
        n = v.shape[0]
        ll += r['ll']
        initC[n] += r['init']
        transC[n] += r['trans']
        phoneC += np.sum(r['cpc'], axis=0)                                    # :231
        cC_all.append(r['cC'])
        pz_all.append(r['pz'])
        ph_all.append(r['ph'])
        cpc_all.append(r['cpc'])
    new = dict(params)
    new['init'], new['trans'] = {}, {}
    for m in lens:
        new['init'][m] = initC[m] / np.sum(initC[m])                          # :237
        tot = np.sum(transC[m], axis=1)                                       # :242
        tr = params['trans'][m].copy()
        for s in range(m):
            if tot[s] != 0:
                tr[s] = transC[m][s] / tot[s]
        new['trans'][m] = tr
    norm = np.sum(np.maximum(phoneC, EPS), axis=-1)                           # :250
    new['phone_probs'] = (phoneC.T / norm).T
    lr, mom = params['lr'], params['momentum']
    D = feats[0].shape[1]
    dWV = np.zeros((K, D + 1))                                                # updateSoftmaxWeightV :503-517
    for v, cC, pz in zip(feats, cC_all, pz_all):
        vc = np.concatenate([v, np.ones((v.shape[0], 1))], axis=1)
        dWV += 1.0 / N * (cC - pz).T @ vc
    new['WV'] = (1.0 - mom) * params['WV'] + lr * dWV
    Da = audio[0].shape[1]
    dWA = np.zeros((nPh, Da + 1))                                             # updateSoftmaxWeightA :527-541
    for a, cpc, ph in zip(audio, cpc_all, ph_all):
        ac = np.concatenate([a, np.ones((a.shape[0], 1))], axis=1)
        Delta = np.sum(cpc, axis=1) - ph
        dWA += 1.0 / N * Delta.T @ ac
    new['WA'] = (1.0 - mom) * params['WA'] + lr * dWA
    if update_lr and (epoch + 1) % 10 == 0:                                   # :259-260
        new['lr'] = lr / 10
    info = dict(avg_ll=ll / N, initC=initC, transC=transC, phoneC=phoneC, cC=cC_all, pz=pz_all, dWA=dWA)
    return new, info


def initial_params(feats, n_words, n_phones, WV, WA, lr=10.0, momentum=0.0, phone_probs=None):
    """initializeModel :93-141 (uniform init / trans / phoneProbs; WV, WA injected)."""
    lens = sorted({v.shape[0] for v in feats})
    return dict(
        init={m: np.ones((m,)) / m for m in lens},
        trans={m: np.ones((m, m)) / m for m in lens},
        phone_probs=(np.ones((n_words, n_phones)) / n_phones) if phone_probs is None
        else np.array(phone_probs, dtype=float),
        WV=np.array(WV, dtype=float), WA=np.array(WA, dtype=float),
        lr=lr, momentum=momentum, toeplitz=len(lens) >= 6,                    # :407
    )


def avg_loglik(feats, audio, params):
    """computeAvgLogLikelihood, :592-602."""
    ll = 0.0
    for v, a in zip(feats, audio):
        pz = ip.posterior_linear(v, params['WV'])
        _, E = emissions(a, params['WA'], params['phone_probs'])
        fwd = ip.forward(pz, np.arange(a.shape[0]), np.ascontiguousarray(E.T), params['init'][v.shape[0]],
                         params['trans'][v.shape[0]])
        ll += ip.pair_loglik(fwd)
    return ll / len(feats)


def align(v, a, params):
    """align, :604-647: EPS score floor and floored alignProbs normaliser."""
    pz = ip.posterior_linear(v, params['WV'])
    _, E = emissions(a, params['WA'], params['phone_probs'])
    n = v.shape[0]
    return ip.align(pz, np.arange(a.shape[0]), np.ascontiguousarray(E.T), params['init'][n], params['trans'][n],
                    floor_norm=True, floor_scores=True)


def cluster(v, a, params, alignment):
    """cluster, :649-661."""
    pz = ip.posterior_linear(v, params['WV'])
    _, E = emissions(a, params['WA'], params['phone_probs'])
    return ip.cluster(pz, np.arange(a.shape[0]), np.ascontiguousarray(E.T), alignment)


# ==============================================================================================
# ImageAudioGaussianHMMWordDiscoverer -- hmm_dnn/image_audio_gaussian_hmm_word_discoverer.py
# RBF posteriors on both sides (:573-595) and NO EPS floor anywhere in the E- or M-step
# (:241-260, :369-371, :414-415, :449-451, :629-631); align keeps its floors (:646-648).
# ==============================================================================================
def emissions_gaussian(a, musA, width, phone_probs):
    """softmaxLayerA :585-595 and probs_x_given_z :304."""
    ph = ip.posterior_gaussian(a, musA, width)
    return ph, (phone_probs @ ph.T).T


def init_counts_raw(fwd, bwd):
    """updateInitialCounts :362-372 -- un-floored."""
    g = fwd * bwd
    return np.sum(g.sum(-1) / g.sum((1, 2))[:, None], axis=0)


def trans_counts_raw(fwd, bwd, pz, E, A, toeplitz):
    """updateTransitionCounts :388-433 -- un-floored normalisation."""
    T, n = fwd.shape[0], fwd.shape[1]
    d = np.diag(A)
    Aoff = A - np.diag(d)
    out = np.zeros((n, n))
    for t in range(T - 1):
        o = E[t + 1]
        xi = np.diag(np.sum(fwd[t] * d[:, None] * o * bwd[t + 1], axis=-1))
        xi = xi + fwd[t].sum(-1)[:, None] * Aoff * np.sum(pz * o * bwd[t + 1], axis=-1)
        xi = xi / np.sum(xi)
        if toeplitz:
            xi = ip.toeplitz_pool(xi)
        out += xi
    return out


def estep_pair_gaussian(v, a, params):
    n, T = v.shape[0], a.shape[0]
    pz = ip.posterior_gaussian(v, params['musV'], params['width'])
    ph, E = emissions_gaussian(a, params['musA'], params['width'], params['phone_probs'])
    x = np.arange(T)
    obs = np.ascontiguousarray(E.T)
    pi, A = params['init'][n], params['trans'][n]
    fwd = ip.forward(pz, x, obs, pi, A)
    bwd = ip.backward(pz, x, obs, A)
    cpc = concept_phone_counts(fwd, bwd, ph)
    return dict(pz=pz, ph=ph, E=E, ll=math.log(float(np.sum(fwd[-1]))),             # :629-631
                init=init_counts_raw(fwd, bwd),
                trans=trans_counts_raw(fwd, bwd, pz, E, A, params['toeplitz']),
                cpc=cpc, cC=ip.concept_counts(pz, x, obs, pi, A))


def em_iteration_gaussian(feats, audio, params):
    """One epoch body of trainUsingEM (:203-268), is_exact = False."""
    K, nPh = params['phone_probs'].shape
    N = len(feats)
    lens = sorted(params['init'].keys())
    initC = {m: np.zeros((m,)) for m in lens}
    transC = {m: np.zeros((m, m)) for m in lens}
    phoneC = np.zeros((K, nPh))
    cC_all, pz_all, ph_all, cpc_all = [], [], [], []
    ll = 0.0
    for v, a in zip(feats, audio):
        r = estep_pair_gaussian(v, a, params)
        n = v.shape[0]
        ll += r['ll']
        initC[n] += r['init']
        transC[n] += r['trans']
        phoneC += np.sum(r['cpc'], axis=0)
        cC_all.append(r['cC'])
        pz_all.append(r['pz'])
        ph_all.append(r['ph'])
        cpc_all.append(r['cpc'])
    new = dict(params)
    new['init'], new['trans'] = {}, {}
    for m in lens:
        new['init'][m] = initC[m] / np.sum(initC[m])                          # :243
        tot = np.sum(transC[m], axis=1)                                       # :248
        tr = params['trans'][m].copy()
        for s in range(m):
            if tot[s] != 0:
                tr[s] = transC[m][s] / tot[s]
        new['trans'][m] = tr
    new['phone_probs'] = (phoneC.T / np.sum(phoneC, axis=-1)).T               # :260-261
    lr, mom, width = params['lr'], params['momentum'], params['width']
    dV = np.zeros_like(params['musV'])                                        # updateSoftmaxWeightV :529-538
    for v, cC, pz in zip(feats, cC_all, pz_all):
        Delta = cC - pz
        dV += 1.0 / (N * width) * (Delta.T @ v - (np.sum(Delta, axis=0) * params['musV'].T).T)
    new['musV'] = (1.0 - mom) * params['musV'] + lr * dV
    dA = np.zeros_like(params['musA'])                                        # updateSoftmaxWeightA :551-558
    for a, cpc, ph in zip(audio, cpc_all, ph_all):
        Delta = np.sum(cpc, axis=1) - ph
        dA += 1.0 / (N * width) * (Delta.T @ a - (np.sum(Delta, axis=0) * params['musA'].T).T)
    new['musA'] = (1.0 - mom) * params['musA'] + lr * dA
    info = dict(avg_ll=ll / N, initC=initC, transC=transC, phoneC=phoneC, cC=cC_all, cpc=cpc_all, ph=ph_all, dA=dA)
    return new, info


def initial_params_gaussian(feats, n_words, n_phones, musV, musA, width=1.0, lr=10.0, momentum=0.0,
                            phone_probs=None):
    lens = sorted({v.shape[0] for v in feats})
    return dict(
        init={m: np.ones((m,)) / m for m in lens},
        trans={m: np.ones((m, m)) / m for m in lens},
        phone_probs=(np.ones((n_words, n_phones)) / n_phones) if phone_probs is None
        else np.array(phone_probs, dtype=float),
        musV=np.array(musV, dtype=float), musA=np.array(musA, dtype=float), width=width,
        lr=lr, momentum=momentum, toeplitz=len(lens) >= 6)


def align_gaussian(v, a, params):
    """align :633-664 (floored like the linear class)."""
    pz = ip.posterior_gaussian(v, params['musV'], params['width'])
    _, E = emissions_gaussian(a, params['musA'], params['width'], params['phone_probs'])
    n = v.shape[0]
    return ip.align(pz, np.arange(a.shape[0]), np.ascontiguousarray(E.T), params['init'][n], params['trans'][n],
                    floor_norm=True, floor_scores=True)


def cluster_gaussian(v, a, params, alignment):
    pz = ip.posterior_gaussian(v, params['musV'], params['width'])
    _, E = emissions_gaussian(a, params['musA'], params['width'], params['phone_probs'])
    return ip.cluster(pz, np.arange(a.shape[0]), np.ascontiguousarray(E.T), alignment)
