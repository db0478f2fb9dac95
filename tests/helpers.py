"""Shared helpers for the parity tests (golden loading, oracle replay)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

IK_CASES = ['tiny_linear', 'tiny_gaussian', 'short_toeplitz_linear', 'short_toeplitz_gaussian',
            'long_floor_linear', 'long_floor_gaussian', 'mixed_linear', 'mixed_gaussian']
IK_TWOLAYER_CASES = ['mixed_twolayer', 'short_twolayer']


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
    if g['kind'] == 'two-layer':
        return orc.initial_params(g['feats_list'], g['K'], g['P'], 'two-layer', W=g['param0'], mus=g['hidden0'], **kw)
    return orc.initial_params(g['feats_list'], g['K'], g['P'], 'gaussian', mus=g['param0'],
                              width=g['width'], **kw)


def write_ik_files(tmp, g):
    """Write a golden case in the reference's on-disk formats; returns (caps, feats, cfg, extra)."""
    import os
    caps = os.path.join(tmp, 'caps.txt')
    with open(caps, 'w') as f:
        for x in g['phones_list']:
            f.write(' '.join('p%d' % p for p in x) + '\n')
    feats = os.path.join(tmp, 'feats.npz')
    np.savez(feats, **{'arr_%d' % i: v for i, v in enumerate(g['feats_list'])})
    cfg = dict(has_null=False, n_words=g['K'], learning_rate=g['lr'], momentum=g['momentum'],
               width=g['width'], feature_dtype='float64')
    extra = {}
    if 'obs0' in g:
        np.save(os.path.join(tmp, 'obs.npy'), g['obs0'])
        extra['obs'] = os.path.join(tmp, 'obs.npy')
    if g['kind'] == 'linear':
        np.savez(os.path.join(tmp, 'w.npz'), weight=g['param0'][:, :-1], bias=g['param0'][:, -1])
        cfg['image_posterior_weights_file'] = os.path.join(tmp, 'w.npz')
    elif g['kind'] == 'two-layer':
        V0, W0 = g['hidden0'], g['param0']
        np.savez(os.path.join(tmp, 'w.npz'), arr_0=V0[:, :-1], arr_1=V0[:, -1], arr_2=W0[:, :-1], arr_3=W0[:, -1])
        cfg['image_posterior_weights_file'] = os.path.join(tmp, 'w.npz')
        cfg['hidden_dim'] = int(V0.shape[0])
    else:
        np.save(os.path.join(tmp, 'mus.npy'), g['param0'])
        cfg['visual_anchor_file'] = os.path.join(tmp, 'mus.npy')
        if 'obs' in extra:
            cfg['obs_prob_file'] = extra['obs']
    return caps, feats, cfg, extra


def make_model(tmp, g):
    import contextlib
    import io
    import os
    caps, feats, cfg, extra = write_ik_files(tmp, g)
    with contextlib.redirect_stdout(io.StringIO()):
        if g['kind'] == 'linear':
            from multimodalworddiscovery_b200.hmm_dnn.image_phone_hmm_word_discoverer import \
                ImagePhoneHMMWordDiscoverer
            m = ImagePhoneHMMWordDiscoverer(caps, feats, cfg, obsProbFile=extra.get('obs'),
                                            modelName=os.path.join(tmp, 'm'))
        elif g['kind'] == 'two-layer':
            from multimodalworddiscovery_b200.hmm_dnn.image_phone_hmm_dnn_word_discoverer import \
                ImagePhoneHMMDNNWordDiscoverer
            m = ImagePhoneHMMDNNWordDiscoverer(caps, feats, cfg, obsProbFile=extra.get('obs'),
                                               modelName=os.path.join(tmp, 'm'))
        else:
            from multimodalworddiscovery_b200.hmm_dnn.image_phone_gaussian_hmm_word_discoverer import \
                ImagePhoneGaussianHMMWordDiscoverer
            m = ImagePhoneGaussianHMMWordDiscoverer(caps, feats, cfg, modelName=os.path.join(tmp, 'm'))
    return m


HMM_CASES = ['flickr60_prob', 'flickr24_log', 'synth_prob', 'synth_log']


def load_hmm(case):
    z = np.load(os.path.join(GOLDEN, 'hmm_%s.npz' % case))
    g = {k: z[k] for k in z.files}
    to, so = g['tgt_off'], g['src_off']
    g['tgt_list'] = [g['tgt'][to[i]:to[i + 1]] for i in range(len(to) - 1)]
    g['src_list'] = [g['src'][so[i]:so[i + 1]] for i in range(len(so) - 1)]
    g['kind'] = str(g['kind'])
    for k in ('Vt', 'Vf', 'n_iter'):
        g[k] = int(g[k])
    return g
