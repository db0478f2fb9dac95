#!/usr/bin/env python
"""Golden vectors for the hmm/ classes from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py hmm

Cases: a prefix of the reference's shipped data/flickr30k/phoneme_level/flickr30k.txt (real
captions: concepts / phones / blank-line blocks) and a seeded synthetic corpus with repeated
concepts and >= 6 distinct sentence lengths.  Sentences are stored as integer ids over sorted
vocabularies; the dict-of-dict obs tables are stored dense with NaN for absent pairs.
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

from make_golden import HERE, REF, load_ref


def read_blocks(path, n_pairs):
    tgt, src = [], []
    with open(path) as f:
        lines = f.read().split('\n')
    i = 0
    while i + 1 < len(lines) and len(tgt) < n_pairs:
        tgt.append(lines[i].split())
        src.append(lines[i + 1].split())
        i += 3
    return tgt, src


def synth_blocks(rng, N, n_lo, n_hi, T_lo, T_hi, Vt, Vf):
    tgt, src = [], []
    for _ in range(N):
        n = int(rng.integers(n_lo, n_hi + 1))
        T = int(rng.integers(T_lo, T_hi + 1))
        tgt.append(['c%d' % c for c in rng.integers(0, Vt, n)])
        src.append(['p%d' % p for p in rng.integers(0, Vf, T)])
    return tgt, src


def write_blocks(path, tgt, src):
    with open(path, 'w') as f:
        for e, s in zip(tgt, src):
            f.write(' '.join(e) + '\n' + ' '.join(s) + '\n\n')


def dense_obs(obs, tv, fv):
    out = np.full((len(tv), len(fv)), np.nan)
    for tw in obs:
        for fw, v in obs[tw].items():
            out[tv[tw], fv[fw]] = v
    return out


def flat(lens, tabs):
    return np.concatenate([np.asarray(tabs[m], dtype=np.float64).ravel() for m in lens])


def run_case(name, kind, tgt, src, n_iter):
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, 'corpus.txt')
        write_blocks(path, tgt, src)
        if kind == 'prob':
            mod = load_ref('hmm/hmm_word_discoverer.py', 'ref_hmm_prob')
            with contextlib.redirect_stdout(io.StringIO()):
                m = mod.HMMWordDiscoverer(path, modelName=os.path.join(tmp, 'm'))
        else:
            mod = load_ref('hmm/audio_hmm_word_discoverer.py', 'ref_hmm_log')
            with contextlib.redirect_stdout(io.StringIO()):
                m = mod.AudioHMMWordDiscoverer(path, modelName=os.path.join(tmp, 'm'))
        tv = {w: i for i, w in enumerate(sorted({w for e in m.tCorpus for w in e}))}
        fv = {w: i for i, w in enumerate(sorted({w for s in m.fCorpus for w in s}))}
        tgt_ids = [np.array([tv[w] for w in e]) for e in m.tCorpus]
        src_ids = [np.array([fv[w] for w in s]) for s in m.fCorpus]
        lens = sorted(m.lenProb)
        out = dict(kind=kind, n_iter=n_iter, Vt=len(tv), Vf=len(fv), lens=np.array(lens),
                   tgt_off=np.cumsum([0] + [len(e) for e in tgt_ids]), tgt=np.concatenate(tgt_ids),
                   src_off=np.cumsum([0] + [len(s) for s in src_ids]), src=np.concatenate(src_ids),
                   tgt_vocab=np.array(sorted(tv)), src_vocab=np.array(sorted(fv)))
        # the reference re-initialises inside trainUsingEM, so per-iteration tables need a hook:
        # run trainUsingEM(k) for k = 1..n_iter from scratch and record the end state of each
        lls = []
        for k in range(1, n_iter + 1):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                m.init = {mm: (np.log(1. / mm) if kind == 'log' else 1. / mm) * np.ones((mm,)) for mm in lens}
                m.trans = {mm: (np.log(1. / mm) if kind == 'log' else 1. / mm) * np.ones((mm, mm)) for mm in lens}
                m.trainUsingEM(k)
            ll_lines = [ln for ln in buf.getvalue().split('\n') if 'Average Log Likelihood' in ln]
            lls.append(float(ll_lines[-1].split(':')[-1]))
            out['init_%d' % (k - 1)] = flat(lens, m.init)
            out['trans_%d' % (k - 1)] = flat(lens, m.trans)
            out['obs_%d' % (k - 1)] = dense_obs(m.obs, tv, fv)
        # prob: printed LL of epoch k-1 is the LL of the parameters ENTERING that epoch;
        # log : printed LL of epoch k-1 is the LL of the parameters AFTER that epoch's M-step
        out['avg_ll'] = np.array(lls)
        with contextlib.redirect_stdout(io.StringIO()):
            out['final_ll'] = m.computeAvgLogLikelihood()
            m.printAlignment(os.path.join(tmp, 'ali'))
        with open(os.path.join(tmp, 'ali.json')) as f:
            ali = json.load(f)
        out['alignment'] = np.concatenate([np.array(a['alignment']) for a in ali])
        out['align_probs'] = np.concatenate([np.array(a['align_probs']).ravel() for a in ali])
        out['fwd0'] = m.forward(m.tCorpus[0], m.fCorpus[0])
        out['bwd0'] = m.backward(m.tCorpus[0], m.fCorpus[0])
    np.savez_compressed(os.path.join(HERE, 'hmm_%s.npz' % name), **out)
    print('wrote hmm_%s.npz avg_ll=%s' % (name, np.array2string(np.array(lls), precision=6)))


def make_hmm():
    flickr = os.path.join(REF, 'data/flickr30k/phoneme_level/flickr30k.txt')
    tgt, src = read_blocks(flickr, 60)
    run_case('flickr60_prob', 'prob', tgt, src, 3)
    tgt, src = read_blocks(flickr, 24)
    run_case('flickr24_log', 'log', tgt, src, 3)
    rng = np.random.default_rng(20261021)
    tgt, src = synth_blocks(rng, 40, 1, 8, 2, 30, Vt=12, Vf=10)
    run_case('synth_prob', 'prob', tgt, src, 3)
    tgt, src = synth_blocks(rng, 20, 1, 7, 2, 24, Vt=9, Vf=8)
    run_case('synth_log', 'log', tgt, src, 3)


if __name__ == '__main__':
    make_hmm()
