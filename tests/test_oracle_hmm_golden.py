"""Pin the plain-state HMM oracle (oracle/plain_hmm.py) against the reference's hmm/ classes."""
import numpy as np
import pytest

from helpers import HMM_CASES, flatten_tables, load_hmm
from oracle import plain_hmm as ph

RTOL = 1e-9


@pytest.mark.parametrize('case', HMM_CASES)
def test_hmm_oracle_matches_reference(case):
    g = load_hmm(case)
    tgt, src, Vt, Vf = g['tgt_list'], g['src_list'], g['Vt'], g['Vf']
    lens = [int(m) for m in g['lens']]
    log = g['kind'] == 'log'
    if log:
        p = dict(init={m: np.log(1. / m) * np.ones(m) for m in lens},
                 trans={m: np.log(1. / m) * np.ones((m, m)) for m in lens},
                 obs=ph.log_initial_obs(tgt, src, Vt, Vf))
        acc = ph.LogAccumulators(lens, Vt, Vf)
    else:
        p = dict(init={m: np.ones(m) / m for m in lens}, trans={m: np.ones((m, m)) / m for m in lens},
                 obs=ph.prob_initial_obs(tgt, src, Vt, Vf))
    for it in range(g['n_iter']):
        if log:
            p, info = ph.log_em_iteration(tgt, src, p, acc)
        else:
            p, info = ph.prob_em_iteration(tgt, src, p)
        np.testing.assert_allclose(info['avg_ll'], g['avg_ll'][it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(lens, p['init']), g['init_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(lens, p['trans']), g['trans_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(p['obs'], g['obs_%d' % it], rtol=RTOL, equal_nan=True)
    ali, ap = [], []
    for e, f in zip(tgt, src):
        n = len(e)
        fn = ph.log_align if log else ph.prob_align
        path, probs = fn(e, f, p['obs'], p['init'][n], p['trans'][n])
        ali += path
        ap += np.array(probs).ravel().tolist()
    assert np.array_equal(np.array(ali), g['alignment'])
    np.testing.assert_allclose(np.array(ap), g['align_probs'], rtol=1e-8)
    e, f = tgt[0], src[0]
    n = len(e)
    if log:
        np.testing.assert_allclose(ph.log_forward(e, f, p['obs'], p['init'][n], p['trans'][n]), g['fwd0'], rtol=RTOL)
        np.testing.assert_allclose(ph.log_backward(e, f, p['obs'], p['trans'][n]), g['bwd0'], rtol=RTOL)
    else:
        np.testing.assert_allclose(ph.prob_forward(e, f, p['obs'], p['init'][n], p['trans'][n]), g['fwd0'], rtol=RTOL)
        np.testing.assert_allclose(ph.prob_backward(e, f, p['obs'], p['trans'][n]), g['bwd0'], rtol=RTOL)


def test_flickr_prefix_known_answer():
    """SURVEY 8c KAT: first 200 pairs of the shipped flickr30k.txt give epoch-0 LL
    -157.7824402151807 under HMMWordDiscoverer; the 60-pair golden prefix is its sibling."""
    g = load_hmm('flickr60_prob')
    assert g['avg_ll'][0] == pytest.approx(-151.312323, abs=1e-5)
