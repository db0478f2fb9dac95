#!/usr/bin/env python
"""Pin the shared-memory layouts and descriptor fields of tcgen05.mma kind::tf32 on the GPU box: build operand images
in NumPy for candidate layouts, run ONE MMA (mwd_umma_probe) and compare the accumulator with A @ B.T.
    python tools/umma_layout_probe.py          (needs a B200)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalworddiscovery_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device('cuda', 0)
SW_NONE, SW128, SW64, SW32, SW128_32B = 0, 2, 4, 6, 1


def swz(off, kind):
    if kind == SW128:
        return off ^ (((off >> 7) & 7) << 4)
    if kind == SW64:
        return off ^ (((off >> 7) & 3) << 4)
    if kind == SW32:
        return off ^ (((off >> 7) & 1) << 4)
    if kind == SW128_32B:
        return off ^ (((off >> 7) & 3) << 5)
    return off


def image(X, offset_fn, kind, nbytes):
    img = np.zeros(nbytes // 4, dtype=np.float32)
    for mn in range(X.shape[0]):
        for k in range(X.shape[1]):
            o = swz(offset_fn(mn, k), kind)
            img[o // 4] = X[mn, k]
    return img


def desc(lbo, sbo, kind):
    return ((lbo >> 4) & 0x3fff) << 16 | ((sbo >> 4) & 0x3fff) << 32 | 1 << 46 | kind << 61


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run(a_img, b_img, ad, bd, idsc, N, n_mma=1, a_step=0, b_step=0):
    a = torch.from_numpy(a_img.view(np.uint32).astype(np.int64).astype(np.uint32).view(np.int32).copy()).to(dev)
    b = torch.from_numpy(b_img.view(np.int32).copy()).to(dev)
    out = torch.full((128, N), -77.0, dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mwd_umma_probe(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), ad, bd, idsc, N, n_mma, a_step,
                                  b_step, out.data_ptr(), st))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def report(name, got, want):
    err = np.abs(got - want).max()
    print('%-70s max|err| %-10.3g %s' % (name, err, 'OK' if err < 1e-3 else ('ALL-ZERO' if not got.any() else 'wrong')))


rng = np.random.default_rng(0)
M, N, K = 128, 16, 8
A = rng.integers(-8, 9, (M, K)).astype(np.float32)
B = rng.integers(-8, 9, (N, K)).astype(np.float32)
want = A @ B.T

# operand images -----------------------------------------------------------------------------------------------
kmaj_sw128 = lambda sbo: (lambda mn, k: (mn // 8) * sbo + (mn % 8) * 128 + k * 4)
kmaj_none = lambda lbo, sbo: (lambda mn, k: (mn // 8) * sbo + (k // 4) * lbo + (mn % 8) * 16 + (k % 4) * 4)
mnmaj_sw128 = lambda lbo, sbo: (lambda mn, k: (mn // 32) * lbo + (k // 8) * sbo + (k % 8) * 128 + (mn % 32) * 4)
mnmaj_none = lambda lbo, sbo: (lambda mn, k: (mn // 4) * sbo + (k // 8) * lbo + (k % 8) * 16 + (mn % 4) * 4)
mnmaj_sw128_32b = lambda lbo, sbo: (lambda mn, k: (mn // 32) * lbo + (k // 4) * sbo + (k % 4) * 128 + (mn % 32) * 4)

A_k128 = image(A, kmaj_sw128(1024), SW128, 32768)
B_k128 = image(B, kmaj_sw128(1024), SW128, 8192)
report('A K-major SW128 / B K-major SW128 (the posterior kernel)', run(A_k128, B_k128, desc(16, 1024, SW128), desc(16, 1024, SW128), idesc(M, N, 0, 0), N), want)

for lbo, sbo in [(128, 256), (256, 128)]:
    B_kn = image(B, kmaj_none(lbo, sbo), SW_NONE, 8192)
    report('B K-major no-swizzle, image LBO=%d SBO=%d, desc same' % (lbo, sbo), run(A_k128, B_kn, desc(16, 1024, SW128), desc(lbo, sbo, SW_NONE), idesc(M, N, 0, 0), N), want)
    report('B K-major no-swizzle, image LBO=%d SBO=%d, desc swapped' % (lbo, sbo), run(A_k128, B_kn, desc(16, 1024, SW128), desc(sbo, lbo, SW_NONE), idesc(M, N, 0, 0), N), want)

for lbo, sbo in [(2048, 1024), (4096, 1024)]:
    A_mn = image(A, mnmaj_sw128(lbo, sbo), SW128, 32768)
    report('A MN-major SW128, image LBO=%d SBO=%d, desc same' % (lbo, sbo), run(A_mn, B_k128, desc(lbo, sbo, SW128), desc(16, 1024, SW128), idesc(M, N, 1, 0), N), want)
    report('A MN-major SW128, image LBO=%d SBO=%d, desc swapped' % (lbo, sbo), run(A_mn, B_k128, desc(sbo, lbo, SW128), desc(16, 1024, SW128), idesc(M, N, 1, 0), N), want)

for lbo, sbo in [(2048, 512), (4096, 512)]:
    A_mn = image(A, mnmaj_sw128_32b(lbo, sbo), SW128_32B, 32768)
    report('A MN-major SW128_BASE32B, image LBO=%d SBO=%d, desc same' % (lbo, sbo), run(A_mn, B_k128, desc(lbo, sbo, SW128_32B), desc(16, 1024, SW128), idesc(M, N, 1, 0), N), want)
    report('A MN-major SW128_BASE32B, image LBO=%d SBO=%d, desc swapped' % (lbo, sbo), run(A_mn, B_k128, desc(sbo, lbo, SW128_32B), desc(16, 1024, SW128), idesc(M, N, 1, 0), N), want)

for lbo, sbo in [(128, 2048), (2048, 128)]:
    A_mn = image(A, mnmaj_none(lbo, sbo), SW_NONE, 96 * 1024)
    report('A MN-major no-swizzle, image LBO=%d SBO=%d, desc same' % (lbo, sbo), run(A_mn, B_k128, desc(lbo, sbo, SW_NONE), desc(16, 1024, SW128), idesc(M, N, 1, 0), N), want)
    report('A MN-major no-swizzle, image LBO=%d SBO=%d, desc swapped' % (lbo, sbo), run(A_mn, B_k128, desc(sbo, lbo, SW_NONE), desc(16, 1024, SW128), idesc(M, N, 1, 0), N), want)

# B MN-major (N-major) variants, for completeness
Bw = rng.integers(-8, 9, (32, K)).astype(np.float32)
want32 = A @ Bw.T
B_mn = image(Bw, mnmaj_sw128(1024, 1024), SW128, 8192)
report('B MN-major SW128 (N=32)', run(A_k128, B_mn, desc(16, 1024, SW128), desc(1024, 1024, SW128), idesc(M, 32, 0, 1), 32), want32)

# two k-steps through descriptor advance (K-major SW128: +32 bytes = +2)
A2 = rng.integers(-8, 9, (M, 16)).astype(np.float32)
B2 = rng.integers(-8, 9, (N, 16)).astype(np.float32)
report('K-major SW128, 2 MMAs, desc += 2', run(image(A2, kmaj_sw128(1024), SW128, 32768), image(B2, kmaj_sw128(1024), SW128, 8192), desc(16, 1024, SW128), desc(16, 1024, SW128), idesc(M, N, 0, 0), N, 2, 2, 2), A2 @ B2.T)
A_mn2 = image(A2, mnmaj_sw128_32b(2048, 512), SW128_32B, 32768)
report('A MN-major SW128_BASE32B, 2 MMAs, A desc += 64 (1024 B), B K-major SW128 += 2', run(A_mn2, image(B2, kmaj_sw128(1024), SW128, 8192), desc(2048, 512, SW128_32B), desc(16, 1024, SW128), idesc(M, N, 1, 0), N, 2, 64, 2), A2 @ B2.T)
B_kn2 = image(B2, lambda mn, k: (k // 8) * 512 + kmaj_none(128, 256)(mn, k % 8), SW_NONE, 8192)
report('A MN-major SW128_BASE32B + B K-major no-swizzle (the gradient kernel), 2 MMAs', run(A_mn2, B_kn2, desc(2048, 512, SW128_32B), desc(128, 256, SW_NONE), idesc(M, N, 1, 0), N, 2, 64, 32), A2 @ B2.T)
