#!/usr/bin/env python
"""Throughput of one EM iteration of the plain-state hmm/ classes (HMMWordDiscoverer = prob domain,
AudioHMMWordDiscoverer = log domain) on one B200: synthetic Flickr30k-shaped pairs
(n ~ empirical 1..8 concept states from a 1546-word vocabulary, T ~ clip(N(49,13),15,125) phones of 69
types).  Prints one JSON line per class: device-timed ms per iteration (CUDA events, corpus resident).

    python profiles/scripts/bench_hmm.py [--pairs 200000] [--steps 5] [--warmup 2]
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
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--zipf', type=float, default=0.0,
                    help='>0: concept and phone ids drawn from Zipf(s) unigrams instead of uniformly')
    ap.add_argument('--domain', choices=('both', 'prob', 'log'), default='both')
    args = ap.parse_args()
    import torch
    from multimodalworddiscovery_b200.engine_hmm import PackedSentences, PlainHMMEngine
    Vt, Vf = 1546, 69
    rng = np.random.default_rng(20261018)
    pmf = np.array([0.12, 0.28, 0.27, 0.17, 0.09, 0.04, 0.02, 0.01])
    ns = rng.choice(8, size=args.pairs, p=pmf) + 1
    Ts = np.clip(np.round(rng.normal(49, 13, args.pairs)), 15, 125).astype(np.int64)
    if args.zipf > 0:
        pt = 1.0 / np.arange(1, Vt + 1) ** args.zipf
        pf = 1.0 / np.arange(1, Vf + 1) ** args.zipf
        all_t = rng.choice(Vt, size=int(ns.sum()), p=pt / pt.sum())
        all_f = rng.choice(Vf, size=int(Ts.sum()), p=pf / pf.sum())
        tgt = np.split(all_t, np.cumsum(ns)[:-1])
        src = np.split(all_f, np.cumsum(Ts)[:-1])
    else:
        tgt = [rng.integers(0, Vt, n) for n in ns]
        src = [rng.integers(0, Vf, T) for T in Ts]
    pk = PackedSentences(tgt, src, Vf)
    for log_domain in {'both': (False, True), 'prob': (False,), 'log': (True,)}[args.domain]:
        eng = PlainHMMEngine(pk, Vt, Vf, log_domain)
        lens = pk.lens
        if log_domain:
            init = {m: np.log(np.ones(m) / m) for m in lens}
            trans = {m: np.log(np.ones((m, m)) / m) for m in lens}
            obs = np.log(np.ones((Vt, Vf)) / Vf)
        else:
            init = {m: np.ones(m) / m for m in lens}
            trans = {m: np.ones((m, m)) / m for m in lens}
            obs = np.ones((Vt, Vf)) / Vf
        eng.set_params(init, trans, obs)
        snap = [t.clone() for t in (eng.init_t, eng.trans_t, eng.obs)]

        def step():
            for dst, s in zip((eng.init_t, eng.trans_t, eng.obs), snap):
                dst.copy_(s)
            if log_domain:
                eng.reset_accumulators()
            return eng.em_iteration()

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
        eng.align()          # first call loads the decode kernels
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            eng.align()
        e1.record()
        torch.cuda.synchronize()
        ms_align = e0.elapsed_time(e1) / args.steps
        print(json.dumps({'metric': 'em_caption_pairs_per_sec',
                          'class': 'AudioHMMWordDiscoverer (log domain)' if log_domain else 'HMMWordDiscoverer (prob domain)',
                          'pairs': args.pairs, 'slots': int(pk.slot_off[-1]), 'ms_per_step': ms,
                          'value': args.pairs / (ms * 1e-3), 'unit': 'pairs/s', 'align_ms_per_pass': ms_align,
                          'align_pairs_per_sec': args.pairs / (ms_align * 1e-3),
                          'avg_log_likelihood': float(ll) / args.pairs, 'n_gpus': 1, 'dtype': 'f64', 'data': 'synthetic'}))
        del eng
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
