"""Shared helpers for the parity tests (golden loading, oracle replay)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

IK_CASES = ['tiny_linear', 'tiny_gaussian', 'short_toeplitz_linear', 'short_toeplitz_gaussian',
            'long_floor_linear', 'long_floor_gaussian', 'mixed_linear', 'mixed_gaussian']


def load_ik(case):
    z = np.load(os.path.join(GOLDEN, 'ik_%s.npz' % case))
    g = {k: z[k] for k in z.files}
    fo, po = g['feat_off'], g['phone_off']
    g['feats_list'] = [g['feats'][fo[i]:fo[i + 1]] for i in range(len(fo) - 1)]
    g['phones_list'] = [g['phones'][po[i]:po[i + 1]] for i in range(len(po) - 1)]
    g['kind'] = str(g['kind'])
    for k in ('K', 'P', 'D', 'n_iter'):
        g[k] = int(g[k])
    for k in ('lr', 'momentum', 'width'):
        g[k] = float(g[k])
    return g


def flatten_tables(lens, tabs):
    return np.concatenate([np.asarray(tabs[int(m)], dtype=np.float64).ravel() for m in lens])


def oracle_params_from_golden(g):
    from oracle import image_phone_hmm as orc
    kw = dict(lr=g['lr'], momentum=g['momentum'], obs=g.get('obs0'))
    if g['kind'] == 'linear':
        return orc.initial_params(g['feats_list'], g['K'], g['P'], 'linear', W=g['param0'], **kw)
    return orc.initial_params(g['feats_list'], g['K'], g['P'], 'gaussian', mus=g['param0'],
                              width=g['width'], **kw)
