"""Pin the NumPy oracle against vectors produced by the unmodified reference classes
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from helpers import IK_CASES, IK_TWOLAYER_CASES, flatten_tables, load_ik, oracle_params_from_golden
from oracle import image_phone_hmm as orc

RTOL = 1e-9


@pytest.mark.parametrize('case', IK_CASES)
def test_ik_oracle_matches_reference(case):
    g = load_ik(case)
    p = oracle_params_from_golden(g)
    feats, phones = g['feats_list'], g['phones_list']
    key = 'W' if g['kind'] == 'linear' else 'mus'
    for it in range(g['n_iter']):
        p, info = orc.em_iteration(feats, phones, p, g['kind'])
        np.testing.assert_allclose(info['avg_ll'], g['avg_ll'][it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(g['lens'], p['init']), g['init_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(g['lens'], p['trans']), g['trans_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(p['obs'], g['obs_%d' % it], rtol=RTOL, atol=0)
        np.testing.assert_allclose(p[key], g['param_%d' % it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(np.concatenate(info['cC']), g['cC_%d' % it], rtol=RTOL, atol=1e-300)
        np.testing.assert_allclose(np.concatenate(info['cA']), g['cA_%d' % it], rtol=RTOL, atol=1e-300)
    # decode with the final parameters and the last E-step's conceptCountsA
    ali, ic, ca, ap = [], [], [], []
    for ex, (v, x) in enumerate(zip(feats, phones)):
        n = v.shape[0]
        pz = orc.posterior(v, p, g['kind'])
        path, probs = orc.align(pz, x, p['obs'], p['init'][n], p['trans'][n],
                                floor_norm=(g['kind'] == 'gaussian'))
        cl, _ = orc.cluster(pz, x, p['obs'], path)
        ali += path
        ic += cl
        ca += np.argmax(info['cA'][ex], axis=1).tolist()
        ap += np.array(probs).ravel().tolist()
    assert np.array_equal(np.array(ali), g['alignment'])
    assert np.array_equal(np.array(ic), g['image_concepts'])
    assert np.array_equal(np.array(ca), g['concept_alignment'])
    np.testing.assert_allclose(np.array(ap), g['align_probs'], rtol=1e-8)
    # dense forward/backward of pair 0
    v, x = feats[0], phones[0]
    n = v.shape[0]
    pz = orc.posterior(v, p, g['kind'])
    np.testing.assert_allclose(orc.forward(pz, x, p['obs'], p['init'][n], p['trans'][n]), g['fwd0'], rtol=RTOL)
    np.testing.assert_allclose(orc.backward(pz, x, p['obs'], p['trans'][n]), g['bwd0'], rtol=RTOL)


def test_floor_known_answer():
    """Self-derived KAT: if every raw likelihood is < 1e-50 the epoch-0 average
    log-likelihood is exactly log(1e-50)."""
    g = load_ik('long_floor_linear')
    assert g['avg_ll'][0] == pytest.approx(-115.12925464970229, abs=0, rel=1e-15)


@pytest.mark.parametrize('case', IK_TWOLAYER_CASES)
def test_twolayer_oracle_matches_reference(case):
    """ImagePhoneHMMDNNWordDiscoverer (ReLU-MLP posterior, backprop M-step, floored init/trans,
    un-floored Viterbi scores) -- SURVEY 8(f1)."""
    g = load_ik(case)
    p = oracle_params_from_golden(g)
    feats, phones = g['feats_list'], g['phones_list']
    for it in range(g['n_iter']):
        p, info = orc.em_iteration(feats, phones, p, 'two-layer')
        np.testing.assert_allclose(info['avg_ll'], g['avg_ll'][it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(g['lens'], p['init']), g['init_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(g['lens'], p['trans']), g['trans_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(p['obs'], g['obs_%d' % it], rtol=RTOL, atol=0)
        np.testing.assert_allclose(p['W'], g['param_%d' % it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(p['V'], g['hidden_%d' % it], rtol=1e-8, atol=1e-12)
    ali, ic, ap, cp = [], [], [], []
    for v, x in zip(feats, phones):
        n = v.shape[0]
        pz = orc.posterior(v, p, 'two-layer')
        path, probs = orc.align(pz, x, p['obs'], p['init'][n], p['trans'][n], floor_norm=True, floor_scores=False)
        cl, sc = orc.cluster(pz, x, p['obs'], path)
        ali += path
        ic += cl
        ap += np.array(probs).ravel().tolist()
        cp += np.array(sc).ravel().tolist()
    assert np.array_equal(np.array(ali), g['alignment'])
    assert np.array_equal(np.array(ic), g['image_concepts'])
    np.testing.assert_allclose(np.array(ap), g['align_probs'], rtol=1e-8)
    np.testing.assert_allclose(np.array(cp), g['cluster_probs'], rtol=1e-8, atol=1e-300)
