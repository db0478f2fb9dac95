"""CUDA-graph replay of the EM iteration (IKEngine.em_iteration_graph, the small-corpus mode of trainUsingEM): the
same kernels in the same order, so parameters, log-likelihood and decode output must be BITWISE those of the
launch-by-launch iteration -- including across a learning-rate change (a new graph) and a parameter overwrite."""
import numpy as np
import pytest

from helpers import load_ik, oracle_params_from_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case,mixed', [('mixed_linear', 0), ('short_toeplitz_linear', 0), ('mixed_gaussian', 0),
                                        ('mixed_linear', 'all')])
def test_graph_iteration_is_bitwise_the_plain_iteration(case, mixed):
    from multimodalworddiscovery_b200.corpus import pack_pairs
    from multimodalworddiscovery_b200.engine import IKEngine
    g = load_ik(case)
    p = oracle_params_from_golden(g)
    gaussian = g['kind'] == 'gaussian'
    engs = []
    for _ in range(2):
        pk = pack_pairs(g['feats_list'], g['phones_list'], feat_dtype=np.float32)
        e = IKEngine(pk, g['K'], g['P'], gaussian=gaussian, mixed_precision=mixed)
        e.set_params(p['init'], p['trans'], p['obs'], p['mus'] if gaussian else p['W'])
        engs.append(e)
    lr = g['lr']
    for it in range(6):
        a = engs[0].em_iteration(lr, g['momentum'], g['width'], with_cA=False)
        b = engs[1].em_iteration_graph(lr, g['momentum'], g['width'])
        assert float(a) == float(b), it
        for x, y in zip(engs[0].get_params()[2:], engs[1].get_params()[2:]):
            assert np.array_equal(x, y, equal_nan=True), it
        if it == 2:
            lr /= 10          # second graph
        if it == 3:           # parameters replaced from the host between replays
            for e in engs:
                e.set_params(p['init'], p['trans'], p['obs'], p['mus'] if gaussian else p['W'])
    d0 = engs[0].decode(floor_norm=gaussian, want_probs=False, width=g['width'])
    d1 = engs[1].decode(floor_norm=gaussian, want_probs=False, width=g['width'])
    assert (d0[0] == d1[0]).all() and (d0[1] == d1[1]).all()
    ca0, ca1 = engs[0].concept_alignment(), engs[1].concept_alignment()
    assert (ca0 == ca1).all()
