#!/usr/bin/env python
"""Error trajectory of a mixed-precision mode against the float64 path over 20 EM iterations on the acceptance
corpus of tests/test_gpu_mixed_precision.py (20 000 coco5-shaped pairs, K = 65, D = 512).

  python profiles/scripts/mixed_trajectory.py concept all mixed      (one column group per mode)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_mixed_precision import _engine, _synth   # noqa: E402


def main(modes, n_pairs=20000, iters=20):
    rng = np.random.default_rng(2)
    K, P, D = 65, 49, 512
    feats, phones, cent = _synth(rng, n_pairs, K, P, D, [5], 15, 90)
    post = 0.01 * rng.standard_normal((K, D + 1))
    init, trans, obs = {5: np.ones(5) / 5}, {5: np.ones((5, 5)) / 5}, np.ones((K, P)) / P
    engs = []
    for mixed in [0] + list(modes):
        eng = _engine(feats, phones, K, P, False, mixed)
        eng.set_params(init, trans, obs, post)
        engs.append(eng)
    lr = 0.1
    print('# max relative difference to the float64 path (obs: per entry; W: of the table scale; LL relative)')
    print('iter ' + ' '.join('%-38s' % ('%s: obs / W / LL' % m) for m in modes))
    for it in range(iters):
        lls = [float(e.em_iteration(lr, 0.0, 1.0)) for e in engs]
        ref = engs[0].get_params()
        cols = []
        for e, ll in zip(engs[1:], lls[1:]):
            b = e.get_params()
            eo = np.max(np.abs(b[2] - ref[2]) / np.maximum(np.abs(ref[2]), 1e-300))
            ew = np.abs(b[3] - ref[3]).max() / np.abs(ref[3]).max()
            cols.append('%-38s' % ('%.2e / %.2e / %.1e' % (eo, ew, abs(ll - lls[0]) / abs(lls[0]))))
        print('%4d ' % it + ' '.join(cols), flush=True)
        if (it + 1) % 10 == 0:
            lr /= 10


if __name__ == '__main__':
    main(sys.argv[1:] or ['concept'])
