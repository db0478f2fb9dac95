"""GPU parity of the plain-state HMM kernels (hmm/ classes) vs the reference goldens."""
import numpy as np
import pytest

from helpers import HMM_CASES, flatten_tables, load_hmm
from oracle import plain_hmm as ph

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _unsort(pk, arr, off):
    arr = np.asarray(arr)
    out = [None] * pk.n_pairs
    for s, ex in enumerate(pk.order):
        out[ex] = arr[off[s]:off[s + 1]]
    return np.concatenate(out)


@pytest.mark.parametrize('split_postings', [False, True])
@pytest.mark.parametrize('case', HMM_CASES)
def test_hmm_em_matches_reference_golden(case, split_postings, monkeypatch):
    from multimodalworddiscovery_b200.engine_hmm import PackedSentences, PlainHMMEngine
    if split_postings:
        # every (concept, phone) entry with more than 4 postings goes through the grid-wide split
        # reduction that normally only serves the few huge entries of a Zipf-distributed corpus
        monkeypatch.setenv('MWD_HMM_POST_BIG', '4')
    g = load_hmm(case)
    tgt, src, Vt, Vf = g['tgt_list'], g['src_list'], g['Vt'], g['Vf']
    lens = [int(m) for m in g['lens']]
    log = g['kind'] == 'log'
    pk = PackedSentences(tgt, src, Vf)
    eng = PlainHMMEngine(pk, Vt, Vf, log)
    if log:
        init = {m: np.log(1. / m) * np.ones(m) for m in lens}
        trans = {m: np.log(1. / m) * np.ones((m, m)) for m in lens}
        obs0 = ph.log_initial_obs(tgt, src, Vt, Vf)
    else:
        init = {m: np.ones(m) / m for m in lens}
        trans = {m: np.ones((m, m)) / m for m in lens}
        obs0 = ph.prob_initial_obs(tgt, src, Vt, Vf)
    eng.set_params(init, trans, obs0)
    N = len(tgt)
    for it in range(g['n_iter']):
        ll = float(eng.em_iteration()) / N
        if log:
            ll = float(eng.loglik_sum()) / N      # the log class prints the LL AFTER its M-step
        np.testing.assert_allclose(ll, g['avg_ll'][it], rtol=RTOL)
        i2, t2, o2 = eng.get_params()
        np.testing.assert_allclose(flatten_tables(lens, i2), g['init_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(lens, t2), g['trans_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(o2, g['obs_%d' % it], rtol=RTOL, equal_nan=True)
    ali, ap = eng.align()
    assert np.array_equal(_unsort(pk, ali.cpu().numpy(), pk.src_off), g['alignment'])
    np.testing.assert_allclose(_unsort(pk, ap.cpu().numpy(), pk.ap_off), g['align_probs'], rtol=1e-8)
    al, be = eng.dense_sweeps()
    s0 = int(np.flatnonzero(pk.order == 0)[0])
    lo, hi = pk.slot_off[s0], pk.slot_off[s0 + 1]
    np.testing.assert_allclose(al.cpu().numpy()[lo:hi].reshape(g['fwd0'].shape), g['fwd0'], rtol=RTOL)
    np.testing.assert_allclose(be.cpu().numpy()[lo:hi].reshape(g['bwd0'].shape), g['bwd0'], rtol=RTOL)
