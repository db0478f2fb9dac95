#!/usr/bin/env python
"""One EM iteration of the segment-embedding HMM (SegEmbedHMMWordDiscoverer's acoustic
model) on one B200: 6 610 utterances, S ~ clip(N(26,10),4,98) segments of 120-d embeddings, states =
NULL + n concepts (n ~ empirical Flickr 1..8, words Zipf(1.0) over a 1 422-word vocabulary), one Gaussian
per word.  Device-timed (CUDA events), corpus resident.  Times the iteration with the sliced Gaussian
statistics kernel (default) and with the CTA-per-word one (MWD_GAUSS_STATS_SPLIT=0) in the same process.

    python profiles/scripts/bench_segembed.py [--utts 6610] [--steps 5] [--warmup 2]
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
    ap.add_argument('--utts', type=int, default=6610)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    args = ap.parse_args()
    import torch
    from multimodalworddiscovery_b200.engine_hmm import SegmentHMMEngine
    Vt, D, M = 1422, 120, 1
    rng = np.random.default_rng(20261018 + 4)
    pmf = np.array([0.12, 0.28, 0.27, 0.17, 0.09, 0.04, 0.02, 0.01])
    ns = rng.choice(8, size=args.utts, p=pmf)[:, None].ravel() + 1
    Ss = np.clip(np.round(rng.normal(26, 10, args.utts)), 4, 98).astype(np.int64)
    pw = 1.0 / np.arange(1, Vt)
    cent = rng.standard_normal((Vt, D))
    tgt, embs = [], []
    for n, S in zip(ns, Ss):
        e = np.concatenate([[0], 1 + rng.choice(Vt - 1, size=n, p=pw / pw.sum())])
        st = rng.integers(0, len(e), S)
        embs.append((cent[e[st]] + np.sqrt(0.02) * rng.standard_normal((S, D))).astype(np.float32))
        tgt.append(e)
    lens = sorted({len(e) for e in tgt})
    eng = SegmentHMMEngine(tgt, embs, Vt, M)
    eng.set_chain_params({m: np.log(1. / m) * np.ones(m) for m in lens},
                         {m: np.log(1. / m) * np.ones((m, m)) for m in lens})
    eng.set_emission_params(np.zeros((Vt, M)), cent[:, None, :] + 0.1 * rng.standard_normal((Vt, M, D)),
                            0.02 * np.ones((Vt, M, D)))
    out = {'metric': 'em_utterances_per_sec', 'class': 'SegEmbedHMMWordDiscoverer acoustic model (config C4)',
           'utterances': args.utts, 'segments': int(Ss.sum()), 'slots': int(eng.pk.n_slots), 'n_gpus': 1,
           'dtype': 'f64', 'data': 'synthetic'}
    for name, env in (('sliced_stats', None), ('cta_per_word_stats', '0')):
        if env is None:
            os.environ.pop('MWD_GAUSS_STATS_SPLIT', None)
        else:
            os.environ['MWD_GAUSS_STATS_SPLIT'] = env
        for _ in range(args.warmup):
            ll = eng.em_iteration()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ll = eng.em_iteration()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[name] = {'ms_per_step': ms, 'utterances_per_sec': args.utts / (ms * 1e-3),
                     'avg_log_likelihood': float(ll) / args.utts}
    os.environ.pop('MWD_GAUSS_STATS_SPLIT', None)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
