"""Parity at BASELINE.json's full size (configs[4]: 1 M MSCOCO-shaped caption-image pairs) through
size-independent properties of the E-step plus oracle spot checks on pairs sampled from the
1 M-pair run (the oracle cannot run a million pairs, it can run any 48 of them).

Properties (all follow from the reference's per-time-step normalisations,
hmm_dnn/image_phone_hmm_word_discoverer.py:355, :396, :462-465, :529):
  * updateInitialCounts adds a vector that sums to 1 for every t  -> sum(initCounts[n]) == sum of T;
  * updateTransitionCounts adds a matrix that sums to 1 for every t < T-1 -> sum(transCounts[n]) == sum of (T-1);
  * conceptCounts rows are normalised over k -> every row sums to 1;
  * log(EPS) <= per-pair log-likelihood <= 0;
  * the counts of the whole corpus equal the fixed-order sum of the counts of its two halves
    (the data-parallel sharding of DESIGN.md section 6 -- pairs are independent inside an E-step).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PAIRS = int(os.environ.get('MWD_FULL_SIZE_PAIRS', '1000000'))
NMAX = 16


def _shard(torch, dev, rank, world):
    import bench
    from multimodalworddiscovery_b200.corpus import pack_sorted_arrays
    sh = bench.make_shard(torch, dev, N_PAIRS, rank, world, 'coco5')
    pk = pack_sorted_arrays(sh['region_off'].cpu().numpy(), sh['phone_off'].cpu().numpy(),
                            sh['feats'].cpu().numpy(), sh['phones'].cpu().numpy(), lens=sh['lens'],
                            n_pairs_global=N_PAIRS)
    return pk, sh['W'].cpu().numpy()


def _params(K, P, lens, seed=7):
    """Non-uniform tables so that floor and non-floor regimes both occur."""
    rng = np.random.default_rng(seed)
    init = {m: (lambda v: v / v.sum())(rng.random(m) + 0.5) for m in lens}
    trans = {m: (lambda v: v / v.sum(1, keepdims=True))(rng.random((m, m)) + 0.5) for m in lens}
    pw = 1.0 / np.arange(1, P + 1) ** 1.2           # the corpus' own phone unigram (bench.make_shard)
    obs = (pw / pw.sum())[None, :] * (0.5 + rng.random((K, P)))
    obs /= obs.sum(1, keepdims=True)
    return init, trans, obs


def _run(torch, dev, rank, world):
    import bench
    from multimodalworddiscovery_b200.engine import IKEngine
    pk, W = _shard(torch, dev, rank, world)
    eng = IKEngine(pk, bench.K_CONCEPTS, bench.P_PHONES, gaussian=False, device=dev, keep_concept_counts_a=False)
    init, trans, obs = _params(bench.K_CONCEPTS, bench.P_PHONES, pk.lens)
    eng.set_params(init, trans, obs, W)
    eng.estep(1.0, with_cA=False)
    torch.cuda.synchronize()
    return pk, eng, dict(init=init, trans=trans, obs=obs, W=W, toeplitz=False)


def test_full_size_properties_and_oracle_spot_checks():
    import torch
    import bench
    from oracle import image_phone_hmm as orc
    dev = torch.device('cuda', 0)
    K, P = bench.K_CONCEPTS, bench.P_PHONES
    pk, eng, params = _run(torch, dev, 0, 1)
    counts = eng.counts.cpu().numpy()
    pe, ie, te = P * K, (NMAX + 1) * NMAX, (NMAX + 1) * NMAX * NMAX
    initC = counts[pe:pe + ie].reshape(NMAX + 1, NMAX)
    transC = counts[pe + ie:pe + ie + te].reshape(NMAX + 1, NMAX * NMAX)
    T = np.diff(pk.phone_off).astype(np.int64)
    n = np.diff(pk.region_off).astype(np.int64)
    assert np.all(np.isfinite(counts))
    for m in pk.lens:
        sel = n == m
        np.testing.assert_allclose(initC[m].sum(), float(T[sel].sum()), rtol=1e-10)
        np.testing.assert_allclose(transC[m].sum(), float((T[sel] - 1).sum()), rtol=1e-10)
    for m in range(NMAX + 1):
        if m not in pk.lens:
            assert not initC[m].any() and not transC[m].any()
    ll = eng.pair_ll[:pk.n_pairs].cpu().numpy()
    assert np.all(ll <= 0.0) and np.all(ll >= np.log(1e-50) - 1e-12)
    np.testing.assert_allclose(counts[-1], ll.sum(), rtol=1e-12)
    # both regimes of the EPS floors must be present for this to be a meaningful run
    floored = np.isclose(ll, np.log(1e-50), rtol=0, atol=1e-12)
    assert 0 < floored.sum() < pk.n_pairs
    rows = eng.cC[:pk.n_regions].sum(1).cpu().numpy()
    np.testing.assert_allclose(rows, 1.0, rtol=1e-11)

    # oracle spot checks: shortest, longest and random pairs, from both floor regimes
    rng = np.random.default_rng(11)
    idx = np.unique(np.concatenate([[0, 1, pk.n_pairs - 2, pk.n_pairs - 1],
                                    rng.choice(pk.n_pairs, 28, replace=False),
                                    rng.choice(np.flatnonzero(floored), 8, replace=False),
                                    rng.choice(np.flatnonzero(~floored), 8, replace=False)]))
    cC = eng.cC
    oparams = dict(params)
    oparams['init'] = {m: np.asarray(v) for m, v in params['init'].items()}
    for s in idx:
        r0, r1 = int(pk.region_off[s]), int(pk.region_off[s + 1])
        p0, p1 = int(pk.phone_off[s]), int(pk.phone_off[s + 1])
        v = np.asarray(pk.feats[r0:r1], dtype=np.float64)
        x = np.asarray(pk.phones[p0:p1])
        ref = orc.estep_pair(v, x, oparams, 'linear')
        np.testing.assert_allclose(ll[s], ref['ll'], rtol=1e-9)
        np.testing.assert_allclose(cC[r0:r1].cpu().numpy(), ref['cC'], rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(eng.pz[r0:r1].cpu().numpy(), ref['pz'], rtol=1e-9, atol=1e-300)

    # sharding invariance: whole corpus == rank 0 of 2 + rank 1 of 2 (fixed order); the halves are
    # the round-robin deal of the (n, T)-sorted order that corpus.shard_positions prescribes
    from multimodalworddiscovery_b200.corpus import pack_sorted_arrays, shard_positions
    from multimodalworddiscovery_b200.engine import IKEngine
    del eng
    torch.cuda.empty_cache()
    feats = np.asarray(pk.feats).reshape(pk.n_pairs, 5, -1)       # coco5: n == 5 for every pair
    halves = []
    for rank in range(2):
        pos = shard_positions(pk.n_pairs, rank, 2)
        Th = T[pos]
        poff = np.concatenate([[0], np.cumsum(Th)]).astype(np.int32)
        # gather the phones of the selected pairs (vectorised ragged gather)
        starts = pk.phone_off[pos].astype(np.int64)
        take = np.repeat(starts - poff[:-1], Th) + np.arange(int(poff[-1]))
        pkh = pack_sorted_arrays(np.arange(len(pos) + 1, dtype=np.int32) * 5, poff,
                                 np.ascontiguousarray(feats[pos]).reshape(len(pos) * 5, -1),
                                 np.ascontiguousarray(np.asarray(pk.phones)[take]), lens=pk.lens,
                                 n_pairs_global=N_PAIRS)
        e2 = IKEngine(pkh, K, P, gaussian=False, device=dev, keep_concept_counts_a=False)
        e2.set_params(params['init'], params['trans'], params['obs'], params['W'])
        e2.estep(1.0, with_cA=False)
        torch.cuda.synchronize()
        halves.append(e2.reduced.cpu().numpy().copy())
        del e2
        torch.cuda.empty_cache()
    both = halves[0] + halves[1]
    np.testing.assert_allclose(both[:len(counts)], counts, rtol=1e-10, atol=1e-300)
