#!/usr/bin/env python
"""Golden vectors of ``simulatedAnnealing`` (hmm_dnn/image_phone_hmm_word_discoverer.py:159-196 and the
gaussian sibling :156-193) from the UNMODIFIED reference classes.

    python tests/golden/make_golden_sa.py        # build container only (needs /root/reference)

Protocol (replayed by tests/test_gpu_sa.py): construct the model from files, inject W / mus through a
file, then seed BOTH global RNGs (``np.random.seed``, ``random.seed``) right before calling
``simulatedAnnealing(numIterations, T0, stepScale)``.  Stored: the inputs, the energy pairs the reference
prints every outer iteration (``Current and previous energy level:  E1 E0`` -- E0 changing between two
lines IS the accept / reject sequence), the final tables, ``lr`` (mutated by the inner trainUsingEM calls)
and the alignment files of every new minimum.
"""
import contextlib
import io
import json
import os
import random
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import flatten_tables, load_ref, synth_ik_corpus, write_ik_files  # noqa: E402


def run_sa_case(name, kind, feats, phones, K, seed, lr, n_outer, T0, step_scale, w_scale=0.5, width=1.0):
    rng = np.random.default_rng(seed)
    D = feats[0].shape[1]
    with tempfile.TemporaryDirectory() as tmp:
        phones = write_ik_files(tmp, feats, phones)
        P = int(max(x.max() for x in phones)) + 1
        cfg = dict(has_null=False, n_words=K, learning_rate=lr, momentum=0.0, width=width)
        if kind == 'linear':
            mod = load_ref('hmm_dnn/image_phone_hmm_word_discoverer.py', 'ref_sa_linear')
            P0 = w_scale * rng.standard_normal((K, D + 1))
            np.savez(os.path.join(tmp, 'w.npz'), weight=P0[:, :-1], bias=P0[:, -1])
            cfg['image_posterior_weights_file'] = os.path.join(tmp, 'w.npz')
            with contextlib.redirect_stdout(io.StringIO()):
                m = mod.ImagePhoneHMMWordDiscoverer(os.path.join(tmp, 'caps.txt'), os.path.join(tmp, 'feats.npz'), cfg,
                                                    modelName=os.path.join(tmp, 'm'))
        else:
            mod = load_ref('hmm_dnn/image_phone_gaussian_hmm_word_discoverer.py', 'ref_sa_gauss')
            P0 = w_scale * rng.standard_normal((K, D))
            np.save(os.path.join(tmp, 'mus.npy'), P0)
            cfg['visual_anchor_file'] = os.path.join(tmp, 'mus.npy')
            with contextlib.redirect_stdout(io.StringIO()):
                m = mod.ImagePhoneGaussianHMMWordDiscoverer(os.path.join(tmp, 'caps.txt'),
                                                            os.path.join(tmp, 'feats.npz'), cfg,
                                                            modelName=os.path.join(tmp, 'm'))
            assert len(m.vCorpus) == len(feats), 'gaussian reference truncates to 30 pairs'
        buf = io.StringIO()
        np.random.seed(seed)
        random.seed(seed)
        with contextlib.redirect_stdout(buf):
            m.simulatedAnnealing(numIterations=n_outer, T0=T0, stepScale=step_scale)
        log = buf.getvalue()
        energies = [(float(a), float(b)) for a, b in
                    re.findall(r'Current and previous energy level:\s+(\S+)\s+(\S+)', log)]
        assert len(energies) == n_outer
        n_updates = len(re.findall(r'^Update \d+ after', log, flags=re.M))
        lens = sorted(m.lenProb)
        out = dict(kind=kind, K=K, P=P, D=D, lr=lr, width=width, seed=seed, n_outer=n_outer, T0=T0,
                   step_scale=step_scale, param0=P0,
                   feat_off=np.cumsum([0] + [v.shape[0] for v in feats]), feats=np.concatenate(feats, axis=0),
                   phone_off=np.cumsum([0] + [len(x) for x in phones]), phones=np.concatenate(phones),
                   lens=np.array(lens), energies=np.array(energies), n_updates=n_updates,
                   final_init=flatten_tables(lens, m.init), final_trans=flatten_tables(lens, m.trans),
                   final_obs=m.obs.copy(), final_param=(m.mus if kind == 'gaussian' else m.W).copy(),
                   final_lr=m.lr)
        for c in range(1, n_updates + 1):
            with open(os.path.join(tmp, 'm_%d_alignment.json' % c)) as f:
                ali = json.load(f)
            out['alignment_%d' % c] = np.concatenate([np.array(a['alignment']) for a in ali])
            out['image_concepts_%d' % c] = np.concatenate([np.array(a['image_concepts']) for a in ali])
            out['concept_alignment_%d' % c] = np.concatenate([np.array(a['concept_alignment']) for a in ali])
    np.savez_compressed(os.path.join(HERE, 'sa_%s.npz' % name), **out)
    acc = [i == 0 or energies[i][1] != energies[i - 1][1] for i in range(len(energies))]
    print('wrote sa_%s.npz: energies %s, %d new minima, final lr %g' % (name, energies, n_updates, m.lr))


def main():
    rng = np.random.default_rng(20261101)
    f, x = synth_ik_corpus(rng, 14, [1, 2, 3, 4], 3, 12, K=5, P=8, D=5)
    # small jumps / high temperature so that both the accept and the reject branch occur
    run_sa_case('linear', 'linear', f, x, K=5, seed=11, lr=0.1, n_outer=5, T0=0.05, step_scale=0.6)
    run_sa_case('gaussian', 'gaussian', f, x, K=5, seed=12, lr=0.1, n_outer=5, T0=0.05, step_scale=0.4, width=2.0)


if __name__ == '__main__':
    main()
