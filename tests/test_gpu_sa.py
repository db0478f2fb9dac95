"""simulatedAnnealing (hmm_dnn/image_phone_hmm_word_discoverer.py:159-196, gaussian :156-193) of the CUDA
classes against goldens recorded from the unmodified reference (tests/golden/make_golden_sa.py): the
energy printed at every outer iteration, the accept / reject walk it implies, the final tables, the
mutated learning rate and the alignment files of every new minimum.  The model stays on the GPU inside
the loop (device snapshots); the host attributes are compared after the call returns."""
import contextlib
import io
import json
import os
import random
import re

import numpy as np
import pytest

from helpers import GOLDEN, flatten_tables, make_model

pytestmark = pytest.mark.gpu


def load_sa(case):
    z = np.load(os.path.join(GOLDEN, 'sa_%s.npz' % case))
    g = {k: z[k] for k in z.files}
    fo, po = g['feat_off'], g['phone_off']
    g['feats_list'] = [g['feats'][fo[i]:fo[i + 1]] for i in range(len(fo) - 1)]
    g['phones_list'] = [g['phones'][po[i]:po[i + 1]] for i in range(len(po) - 1)]
    g['kind'] = str(g['kind'])
    for k in ('K', 'P', 'D', 'seed', 'n_outer', 'n_updates'):
        g[k] = int(g[k])
    for k in ('lr', 'width', 'T0', 'step_scale', 'final_lr'):
        g[k] = float(g[k])
    g['momentum'] = 0.0
    return g


@pytest.mark.parametrize('case', ['linear', 'gaussian'])
def test_simulated_annealing_matches_reference(case, tmp_path):
    g = load_sa(case)
    tmp = str(tmp_path)
    m = make_model(tmp, g)
    np.random.seed(g['seed'])
    random.seed(g['seed'])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m.simulatedAnnealing(numIterations=g['n_outer'], T0=g['T0'], stepScale=g['step_scale'])
    log = buf.getvalue()
    energies = np.array([(float(a), float(b)) for a, b in
                         re.findall(r'Current and previous energy level:\s+(\S+)\s+(\S+)', log)])
    assert energies.shape == g['energies'].shape
    np.testing.assert_allclose(energies, g['energies'], rtol=1e-9)
    # accept / reject walk: E0 of the next line changes exactly when the jump was accepted
    ref_acc = [bool(g['energies'][i + 1][1] == g['energies'][i][0]) for i in range(g['n_outer'] - 1)]
    got_acc = [acc for (_, _, acc) in m.sa_trace][:-1]
    assert got_acc == ref_acc
    assert len(re.findall(r'^Update \d+ after', log, flags=re.M)) == g['n_updates']
    key = 'W' if g['kind'] == 'linear' else 'mus'
    np.testing.assert_allclose(flatten_tables(g['lens'], m.init), g['final_init'], rtol=1e-8)
    np.testing.assert_allclose(flatten_tables(g['lens'], m.trans), g['final_trans'], rtol=1e-8)
    np.testing.assert_allclose(m.obs, g['final_obs'], rtol=1e-8, atol=1e-300)
    np.testing.assert_allclose(getattr(m, key), g['final_param'], rtol=1e-8, atol=1e-12)
    assert m.lr == pytest.approx(g['final_lr'], rel=1e-12)
    for c in range(1, g['n_updates'] + 1):
        with open(os.path.join(tmp, 'm_%d_alignment.json' % c)) as f:
            ali = json.load(f)
        assert np.array_equal(np.concatenate([a['alignment'] for a in ali]), g['alignment_%d' % c])
        assert np.array_equal(np.concatenate([a['image_concepts'] for a in ali]), g['image_concepts_%d' % c])
        assert np.array_equal(np.concatenate([a['concept_alignment'] for a in ali]), g['concept_alignment_%d' % c])
        assert os.path.exists(os.path.join(tmp, 'm_%d_initialprobs.txt' % c))
