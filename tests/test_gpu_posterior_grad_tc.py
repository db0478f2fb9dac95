"""The tcgen05 / TMA posterior-gradient kernel (mwd_ik_posterior_grad_tc_partial, csrc/posterior_grad_tc.cu) against
the float64 DMMA kernel of the same library and NumPy:  grad = (cC - pz)^T [V, 1].
Tolerance: 1e-5 of the gradient's scale (split-TF32 operands, fp32 accumulation over <= 512 rows)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(feats, cC, pz, split_mode, chunks=1):
    import torch
    from multimodalworddiscovery_b200 import _lib
    lib = _lib.load()
    dev = torch.device('cuda', 0)
    R, D = feats.shape
    K = cC.shape[1]
    assert lib.mwd_posterior_grad_tc_supported(0, D, K) == 1
    f = torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float32)).to(dev)
    c = torch.from_numpy(np.ascontiguousarray(cC)).to(dev)
    z = torch.from_numpy(np.ascontiguousarray(pz)).to(dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = []
    for tc in (True, False):
        n = lib.mwd_posterior_grad_tc_partials_len(K, D) if tc else lib.mwd_outer_grad_partials_len(K, D)
        part = torch.full((n,), 3.0, dtype=torch.float64, device=dev)
        grad = torch.full((K, D + 1), -5.0, dtype=torch.float64, device=dev)
        bounds = np.linspace(0, R, chunks + 1).astype(np.int64)
        for ci in range(chunks):
            lo, hi = int(bounds[ci]), int(bounds[ci + 1])
            p = _lib.IkProblem()
            p.n_regions, p.feat_dim, p.feat_is_f64, p.n_concepts = hi - lo, D, 0, K
            p.feats, p.concept_counts, p.pz = f[lo:hi].data_ptr(), c[lo:hi].data_ptr(), z[lo:hi].data_ptr()
            if tc:
                _lib.check(lib.mwd_ik_posterior_grad_tc_partial(C.byref(p), part.data_ptr(), 1 if ci else 0, split_mode, st))
            else:
                _lib.check(lib.mwd_ik_posterior_grad_partial(C.byref(p), part.data_ptr(), 1 if ci else 0, st))
        if tc:
            _lib.check(lib.mwd_posterior_grad_tc_finish(K, D, part.data_ptr(), grad.data_ptr(), st))
        else:
            _lib.check(lib.mwd_ik_posterior_grad_finish(K, D, part.data_ptr(), grad.data_ptr(), st))
        torch.cuda.synchronize()
        out.append(grad.cpu().numpy())
    return out


@pytest.mark.parametrize('split_mode', [0, 1])
@pytest.mark.parametrize('R,D,K,chunks', [(16, 32, 16, 1), (100, 64, 65, 1), (5000, 512, 65, 1), (3001, 512, 100, 2),
                                          (777, 96, 33, 1), (200000, 512, 65, 3), (1111, 256, 128, 1)])
def test_grad_tc_matches_float64(R, D, K, chunks, split_mode):
    rng = np.random.default_rng(R + D + K)
    cent = 10.0 * rng.standard_normal((K, D))
    feats = (cent[rng.integers(0, K, R)] + rng.standard_normal((R, D))).astype(np.float32)
    pz = rng.random((R, K)) ** 4
    pz /= pz.sum(1, keepdims=True)
    cC = rng.random((R, K)) ** 4
    cC /= cC.sum(1, keepdims=True)
    tc, ref = _run(feats, cC, pz, split_mode, chunks)
    orc = (cC - pz).T @ np.concatenate([feats.astype(np.float64), np.ones((R, 1))], axis=1)
    scale = np.abs(orc).max()
    np.testing.assert_allclose(ref, orc, rtol=1e-9, atol=1e-9 * scale)
    err = np.abs(tc - orc).max() / scale
    print('R=%d D=%d K=%d mode=%d: max err / scale %.3e; bias column %.3e' %
          (R, D, K, split_mode, err, np.abs(tc[:, -1] - orc[:, -1]).max()))
    assert np.all(np.isfinite(tc))
    np.testing.assert_allclose(tc[:, -1], orc[:, -1], rtol=1e-9, atol=1e-10)      # float64 column sums
    assert err < 1e-5
