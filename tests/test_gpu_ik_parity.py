"""GPU parity of the (i,k)-state EM path: CUDA kernels (through the C ABI) vs the golden vectors
produced by the unmodified reference, and vs the NumPy oracle on fresh seeded inputs.

Tolerances (north star): log-likelihood and count/parameter tables within 1e-5 relative (we
assert far tighter, 1e-9, because everything is float64); alignments and argmax assignments
bit-exact."""
import numpy as np
import pytest

from helpers import IK_CASES, flatten_tables, load_ik, oracle_params_from_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _engine_for(g, feat_dtype=np.float64, keep_cA=True):
    from multimodalworddiscovery_b200.corpus import pack_pairs
    from multimodalworddiscovery_b200.engine import IKEngine
    pk = pack_pairs(g['feats_list'], g['phones_list'], feat_dtype=feat_dtype)
    eng = IKEngine(pk, g['K'], g['P'], gaussian=(g['kind'] == 'gaussian'), keep_concept_counts_a=keep_cA)
    p = oracle_params_from_golden(g)
    eng.set_params(p['init'], p['trans'], p['obs'], p['W'] if g['kind'] == 'linear' else p['mus'])
    return pk, eng


def _unsort_rows(pk, arr, off):
    """rows stored in sorted-pair order -> original corpus order"""
    arr = np.asarray(arr)
    out = [None] * pk.n_pairs
    for s, ex in enumerate(pk.order):
        out[ex] = arr[off[s]:off[s + 1]]
    return np.concatenate(out, axis=0)


@pytest.mark.parametrize('case', IK_CASES)
def test_em_matches_reference_golden(case):
    g = load_ik(case)
    pk, eng = _engine_for(g)
    N = len(g['feats_list'])
    for it in range(g['n_iter']):
        ll = eng.em_iteration(g['lr'], g['momentum'], g['width'])
        np.testing.assert_allclose(float(ll) / N, g['avg_ll'][it], rtol=RTOL)
        init, trans, obs, post = eng.get_params()
        np.testing.assert_allclose(flatten_tables(g['lens'], init), g['init_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(g['lens'], trans), g['trans_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(obs, g['obs_%d' % it], rtol=RTOL, atol=0)
        np.testing.assert_allclose(post, g['param_%d' % it], rtol=1e-8, atol=1e-12)
        cC = _unsort_rows(pk, eng.cC.cpu().numpy(), pk.region_off)
        cA = _unsort_rows(pk, eng.cA.cpu().numpy(), pk.phone_off)
        np.testing.assert_allclose(cC, g['cC_%d' % it], rtol=RTOL, atol=1e-300)
        np.testing.assert_allclose(cA, g['cA_%d' % it], rtol=RTOL, atol=1e-300)
    # decode: bit-exact integers
    ali, ic, ap = eng.decode(floor_norm=(g['kind'] == 'gaussian'), want_probs=True, width=g['width'])
    ca = eng.concept_alignment()
    assert np.array_equal(_unsort_rows(pk, ali.cpu().numpy(), pk.phone_off), g['alignment'])
    assert np.array_equal(_unsort_rows(pk, ic.cpu().numpy(), pk.region_off), g['image_concepts'])
    assert np.array_equal(_unsort_rows(pk, ca.cpu().numpy(), pk.phone_off), g['concept_alignment'])
    np.testing.assert_allclose(_unsort_rows(pk, ap.cpu().numpy(), pk.ap_offsets()), g['align_probs'],
                               rtol=1e-8)
    # computeAvgLogLikelihood under the final parameters
    np.testing.assert_allclose(float(eng.loglik_sum(g['width'])) / N, float(g['final_ll']), rtol=RTOL)
    # dense forward / backward of pair 0
    v0, x0 = g['feats_list'][0], g['phones_list'][0]
    pz0 = eng.posterior_rows(v0, g['width'])
    np.testing.assert_allclose(eng.dense_sweep(pz0, x0, backward=False), g['fwd0'], rtol=RTOL)
    np.testing.assert_allclose(eng.dense_sweep(pz0, x0, backward=True), g['bwd0'], rtol=RTOL)


@pytest.mark.parametrize('kind', ['linear', 'gaussian'])
@pytest.mark.parametrize('K,P,D,nmax', [(65, 49, 64, 10), (100, 69, 32, 8), (33, 20, 17, 16)])
def test_em_matches_oracle_random(kind, K, P, D, nmax):
    """Fresh seeded corpus at MSCOCO / Flickr concept counts (exercises the KG=9, 13, 8 kernels,
    n up to 16, >= 6 distinct n) against the oracle."""
    from oracle import image_phone_hmm as orc
    from multimodalworddiscovery_b200.corpus import pack_pairs
    from multimodalworddiscovery_b200.engine import IKEngine
    rng = np.random.default_rng(1234 + K + nmax)
    feats, phones = [], []
    cents = rng.standard_normal((K, D))
    for _ in range(40):
        n = int(rng.integers(1, nmax + 1))
        T = int(rng.integers(1, 40))
        v = cents[rng.integers(0, K, n)] + 0.5 * rng.standard_normal((n, D))
        feats.append(v.astype(np.float32).astype(np.float64))
        phones.append(rng.integers(0, P, T))
    obs0 = rng.random((K, P)) + 0.01
    obs0 /= obs0.sum(1, keepdims=True)
    if kind == 'linear':
        p = orc.initial_params(feats, K, P, 'linear', W=0.3 * rng.standard_normal((K, D + 1)), lr=0.1,
                               obs=obs0)
        post0 = p['W']
    else:
        p = orc.initial_params(feats, K, P, 'gaussian', mus=cents + 0.1 * rng.standard_normal((K, D)),
                               width=3.0, lr=0.1, obs=obs0)
        post0 = p['mus']
    pk = pack_pairs(feats, phones, feat_dtype=np.float32)
    eng = IKEngine(pk, K, P, gaussian=(kind == 'gaussian'))
    eng.set_params(p['init'], p['trans'], p['obs'], post0)
    lens = sorted(p['init'])
    for it in range(2):
        p, info = orc.em_iteration(feats, phones, p, kind)
        ll = eng.em_iteration(0.1, 0.0, 3.0 if kind == 'gaussian' else 1.0)
        np.testing.assert_allclose(float(ll) / len(feats), info['avg_ll'], rtol=RTOL)
        init, trans, obs, post = eng.get_params()
        np.testing.assert_allclose(flatten_tables(lens, init), flatten_tables(lens, p['init']), rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(lens, trans), flatten_tables(lens, p['trans']), rtol=RTOL)
        np.testing.assert_allclose(obs, p['obs'], rtol=RTOL)
        np.testing.assert_allclose(post, p['W' if kind == 'linear' else 'mus'], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(_unsort_rows(pk, eng.cC.cpu().numpy(), pk.region_off),
                                   np.concatenate(info['cC']), rtol=RTOL, atol=1e-300)


@pytest.mark.parametrize('case', ['mixed_linear', 'mixed_gaussian', 'short_toeplitz_linear'])
def test_streamed_iteration_equals_resident(case):
    """em_iteration_streamed (chunked H2D overlapped with the kernels) == em_iteration."""
    import torch
    g = load_ik(case)
    pk, eng = _engine_for(g, keep_cA=False)
    _, eng2 = _engine_for(g, keep_cA=False)
    host = {k: torch.from_numpy(np.ascontiguousarray(getattr(pk, k))).pin_memory()
            for k in ('region_off', 'phone_off', 'feats', 'phones')}
    for it in range(2):
        ll1 = float(eng.em_iteration(g['lr'], g['momentum'], g['width'], with_cA=False))
        eng2.feats.zero_()      # prove the data really comes from the host copy
        eng2.phones.zero_()
        ll2 = float(eng2.em_iteration_streamed(host, g['lr'], g['momentum'], g['width'], n_chunks=5))
        assert ll2 == pytest.approx(ll1, rel=1e-13)
        for a, b in zip(eng.get_params(), eng2.get_params()):
            if isinstance(a, dict):
                for m in a:
                    np.testing.assert_allclose(b[m], a[m], rtol=1e-12)
            else:
                np.testing.assert_allclose(b, a, rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(eng2.cC.cpu().numpy(), eng.cC.cpu().numpy(), rtol=1e-12, atol=1e-300)
