#!/usr/bin/env python
"""Build-container probe (needs /root/reference): pairs/s of ONE EM iteration of the UNMODIFIED reference class
(hmm_dnn/image_phone_hmm_word_discoverer.py, trainUsingEM(1)) next to the NumPy oracle port that bench.py times on the
GPU box (the reference tree does not travel), same corpus, one process, BLAS threads as configured.
    python tools/probe_reference_speed.py [--pairs 200] > profiles/r02_reference_vs_port.json"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=200)
    args = ap.parse_args()
    import bench
    from make_golden import load_ref, write_ik_files
    from oracle import image_phone_hmm as orc
    K, P = bench.K_CONCEPTS, bench.P_PHONES
    feats, phones, W = bench.cpu_sample_numpy(args.pairs, 'coco5')
    out = {'pairs': args.pairs, 'workload': 'coco5 (n = 5, T ~ N(50,10), K = 65, P = 49, D = 512)',
           'cpu_count': os.cpu_count(), 'blas_threads_env': os.environ.get('OPENBLAS_NUM_THREADS')}
    with tempfile.TemporaryDirectory() as tmp:
        phones2 = write_ik_files(tmp, feats, [np.asarray(x, dtype=np.int64) for x in phones])
        np.savez(os.path.join(tmp, 'w.npz'), weight=W[:, :-1], bias=W[:, -1])
        cfg = dict(has_null=False, n_words=K, learning_rate=0.1, momentum=0.0, width=1.0,
                   image_posterior_weights_file=os.path.join(tmp, 'w.npz'))
        mod = load_ref('hmm_dnn/image_phone_hmm_word_discoverer.py', 'ref_ik_linear_probe')
        with contextlib.redirect_stdout(io.StringIO()):
            m = mod.ImagePhoneHMMWordDiscoverer(os.path.join(tmp, 'caps.txt'), os.path.join(tmp, 'feats.npz'), cfg,
                                                modelName=os.path.join(tmp, 'm'))
            t0 = time.perf_counter()
            m.trainUsingEM(1, writeModel=False)
            t_ref = time.perf_counter() - t0
    params = orc.initial_params(feats, K, P, 'linear', W=W, lr=0.1)
    params['toeplitz'] = False
    t0 = time.perf_counter()
    orc.em_iteration(feats, phones, params, 'linear')
    t_port = time.perf_counter() - t0
    out['reference_pairs_per_s'] = args.pairs / t_ref
    out['port_pairs_per_s'] = args.pairs / t_port
    out['port_over_reference'] = t_ref / t_port
    out['note'] = ('trainUsingEM(1) of the reference also evaluates computeAvgLogLikelihood (a second forward pass) '
                   'like every reference epoch; the port fuses it, as the CUDA path does')
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
