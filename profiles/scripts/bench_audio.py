#!/usr/bin/env python
"""Throughput of one EM iteration of the dense-emission class (ImageAudioHMMWordDiscoverer, SURVEY 8
f2) on one B200: synthetic MSCOCO-shaped pairs (n = 5 regions of D = 512, T ~ clip(N(50,10),15,125)
frames of Da = 40-d audio features, K = 65 concepts, 42 hidden phones).  Prints one JSON line with
device-timed ms per iteration (CUDA events, inputs resident in HBM) and the per-kernel split.

    python profiles/scripts/bench_audio.py [--pairs 200000] [--steps 3] [--warmup 2]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=200000)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=2)
    args = ap.parse_args()
    import torch
    from multimodalworddiscovery_b200.corpus import pack_sorted_arrays
    from multimodalworddiscovery_b200.engine_audio import IKAudioEngine
    K, nPh, D, Da, n = 65, 42, 512, 40, 5
    rng = np.random.default_rng(20261018)
    N = args.pairs
    T = np.sort(np.clip(np.round(rng.normal(50, 10, N)), 15, 125).astype(np.int64))
    phone_off = np.concatenate([[0], np.cumsum(T)]).astype(np.int32)
    region_off = (np.arange(N + 1) * n).astype(np.int32)
    feats = (10.0 * rng.standard_normal((K, D)))[rng.integers(0, K, N * n)].astype(np.float32)
    feats += rng.standard_normal(feats.shape, dtype=np.float32)
    Tt = int(phone_off[-1])
    audio = rng.standard_normal((Tt, Da), dtype=np.float32).astype(np.float64)
    pk = pack_sorted_arrays(region_off, phone_off, feats, np.arange(Tt, dtype=np.int32), lens=[n])
    eng = IKAudioEngine(pk, audio, K, nPh)
    init = {n: np.ones(n) / n}
    trans = {n: np.ones((n, n)) / n}
    pp = np.ones((K, nPh)) / nPh
    WV = 0.01 * rng.standard_normal((K, D + 1))
    WA = 0.1 * rng.standard_normal((nPh, Da + 1))
    eng.set_params(init, trans, pp, WV)
    eng.set_audio_param(WA)
    snap = [t.clone() for t in (eng.init_t, eng.trans_t, eng.obsT, eng.post, eng.WA)]

    def step():
        for dst, src in zip((eng.init_t, eng.trans_t, eng.obsT, eng.post, eng.WA), snap):
            dst.copy_(src)
        return eng.em_iteration(0.1, 0.0)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ll = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({'metric': 'em_caption_pairs_per_sec', 'class': 'ImageAudioHMMWordDiscoverer', 'pairs': N,
                      'frames': Tt, 'ms_per_step': ms, 'value': N / (ms * 1e-3), 'unit': 'pairs/s',
                      'avg_log_likelihood': float(ll) / N, 'n_gpus': 1, 'dtype': 'f64', 'data': 'synthetic'}))


if __name__ == '__main__':
    main()
