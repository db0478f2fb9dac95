"""The tcgen05 / TMA image-posterior kernel (mwd_posterior_linear_tc, csrc/posterior_tc.cu: split-TF32 operands,
fp32 TMEM accumulators, float64 recombination + softmax) against the float64 DMMA kernel of the same library and
the NumPy oracle.  Tolerance: posteriors above 1e-12 agree to 3e-6 relative on average and 2e-5 at worst (D = 512);
rows sum to one to 1e-12."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(feats, W, split_mode):
    import torch
    from multimodalworddiscovery_b200 import _lib
    lib = _lib.load()
    dev = torch.device('cuda', 0)
    R, D = feats.shape
    K = W.shape[0]
    assert lib.mwd_posterior_tc_supported(0, D, K) == 1
    f = torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float32)).to(dev)
    w = torch.from_numpy(np.ascontiguousarray(W, dtype=np.float64)).to(dev)
    out_tc = torch.full((R, K), -7.0, dtype=torch.float64, device=dev)
    out_64 = torch.empty((R, K), dtype=torch.float64, device=dev)
    scratch = torch.empty((lib.mwd_posterior_tc_scratch_bytes(K, D) // 4,), dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mwd_posterior_linear_tc(f.data_ptr(), R, D, w.data_ptr(), K, out_tc.data_ptr(),
                                           scratch.data_ptr(), split_mode, st))
    _lib.check(lib.mwd_posterior_linear(f.data_ptr(), 0, R, D, w.data_ptr(), K, out_64.data_ptr(), st))
    torch.cuda.synchronize()
    return out_tc.cpu().numpy(), out_64.cpu().numpy()


def _oracle(feats, W):
    x = feats.astype(np.float64) @ W[:, :-1].T + W[:, -1]
    x -= x.max(1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(1, keepdims=True)


@pytest.mark.parametrize('split_mode', [0, 1])
@pytest.mark.parametrize('R,D,K', [(128, 32, 16), (300, 64, 65), (5000, 512, 65), (1000, 512, 100), (777, 96, 33),
                                   (40000, 512, 65), (1111, 512, 128)])
def test_posterior_tc_matches_float64(R, D, K, split_mode):
    rng = np.random.default_rng(R + D + K)
    cent = 10.0 * rng.standard_normal((K, D))
    feats = (cent[rng.integers(0, K, R)] + rng.standard_normal((R, D))).astype(np.float32)
    W = np.concatenate([0.01 * rng.standard_normal((K, D)), rng.standard_normal((K, 1))], axis=1)
    tc, ref = _run(feats, W, split_mode)
    orc = _oracle(feats, W)
    np.testing.assert_allclose(ref, orc, rtol=1e-9, atol=1e-300)
    assert np.all(np.isfinite(tc))
    np.testing.assert_allclose(tc.sum(1), 1.0, rtol=1e-12)
    big = orc > 1e-12
    rel = np.abs(tc - orc)[big] / orc[big]
    print('R=%d D=%d K=%d mode=%d: max rel err %.3e, mean %.3e' % (R, D, K, split_mode, rel.max(), rel.mean()))
    # the fp32 TMEM accumulate truncates: the error grows with D / chunks (see csrc/posterior_tc.cu)
    assert rel.mean() < 3e-6 and rel.max() < 2e-5
    np.testing.assert_allclose(tc, orc, atol=1e-9, rtol=2e-5)
