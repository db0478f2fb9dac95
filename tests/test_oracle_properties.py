"""Size-independent properties of the oracle itself (CPU): they hold for the reference's algorithm by
construction (hmm_dnn/image_phone_hmm_word_discoverer.py:276-465) and pin the restatement beyond the
golden vectors -- the GPU full-size test relies on the same identities."""
import numpy as np
import pytest

from oracle import image_phone_hmm as orc


def _random_pair(rng, n, T, K, P, D):
    v = rng.standard_normal((n, D))
    x = rng.integers(0, P, T)
    W = 0.4 * rng.standard_normal((K, D + 1))
    obs = rng.random((K, P)) + 0.05
    obs /= obs.sum(1, keepdims=True)
    pi = rng.random(n) + 0.2
    pi /= pi.sum()
    A = rng.random((n, n)) + 0.2
    A /= A.sum(1, keepdims=True)
    return v, x, W, obs, pi, A


@pytest.mark.parametrize('seed', range(6))
def test_forward_backward_identities(seed):
    rng = np.random.default_rng(seed)
    n, T, K, P, D = int(rng.integers(1, 7)), int(rng.integers(1, 25)), 9, 6, 5
    v, x, W, obs, pi, A = _random_pair(rng, n, T, K, P, D)
    pz = orc.posterior_linear(v, W)
    np.testing.assert_allclose(pz.sum(1), 1.0, rtol=1e-12)
    fwd = orc.forward(pz, x, obs, pi, A)
    bwd = orc.backward(pz, x, obs, A)
    # sum_{i,k} alpha_t beta_t is the sentence likelihood at EVERY t (the E-step kernels use one normaliser)
    L = fwd[-1].sum()
    np.testing.assert_allclose((fwd * bwd).sum((1, 2)), L, rtol=1e-10)
    assert L > 1e-50                                   # short pairs: above the EPS floor
    # occupancy counts sum to T, transition counts to T-1, state posteriors to 1 per t
    np.testing.assert_allclose(orc.init_counts(fwd, bwd).sum(), T, rtol=1e-10)
    tc = orc.trans_counts(fwd, bwd, pz, x, obs, A, toeplitz=False)
    np.testing.assert_allclose(tc.sum(), T - 1, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(orc.state_counts(fwd, bwd).sum((1, 2)), 1.0, rtol=1e-10)
    # Toeplitz pooling is linear and preserves ... the diagonal sums: total grows by the pooling factor
    tp = orc.trans_counts(fwd, bwd, pz, x, obs, A, toeplitz=True)
    assert tp.shape == (n, n) and np.all(np.isfinite(tp))
    # restricted-chain concept posteriors are distributions over k
    cC = orc.concept_counts(pz, x, obs, pi, A)
    np.testing.assert_allclose(cC.sum(1), 1.0, rtol=1e-10)
    assert np.all(cC >= 0)


def test_floor_regime_is_uniform():
    """Likelihood below EPS: log-likelihood is exactly log(EPS) and the floored occupancy counts are uniform
    (T / n per region) -- SURVEY 8 a6 / a7."""
    rng = np.random.default_rng(3)
    n, T, K, P, D = 4, 160, 9, 30, 5
    v, x, W, obs, pi, A = _random_pair(rng, n, T, K, P, D)
    pz = orc.posterior_linear(v, W)
    fwd = orc.forward(pz, x, obs, pi, A)
    bwd = orc.backward(pz, x, obs, A)
    assert fwd[-1].sum() < 1e-50
    assert orc.pair_loglik(fwd) == np.log(1e-50)
    # steps whose alpha*beta entries are all below EPS contribute exactly 1/n each
    ic = orc.init_counts(fwd, bwd)
    np.testing.assert_allclose(ic.sum(), T, rtol=1e-10)
    small = np.all(fwd * bwd < 1e-50, axis=(1, 2))
    assert small.any()


def test_viterbi_path_is_optimal_on_small_case():
    """align(): the returned path maximises the product of (floored) scores among all region sequences
    of a short caption (brute force)."""
    import itertools
    rng = np.random.default_rng(8)
    n, T, K, P, D = 3, 5, 7, 5, 4
    v, x, W, obs, pi, A = _random_pair(rng, n, T, K, P, D)
    pz = orc.posterior_linear(v, W)
    path, probs = orc.align(pz, x, obs, pi, A)
    p = (pz @ obs[:, x]).T                               # (T, n) marginal emissions

    def score(q):
        s = pi[q[0]] * p[0, q[0]]
        for t in range(1, T):
            s *= A[q[t - 1], q[t]] * p[t, q[t]]
        return s
    best = max(itertools.product(range(n), repeat=T), key=score)
    assert score(tuple(path)) == pytest.approx(score(best), rel=1e-12)
    assert len(probs) == T and len(probs[0]) == n
