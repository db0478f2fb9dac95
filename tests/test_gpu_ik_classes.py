"""The reference-facing classes (same names / signatures / files) against the reference's golden
outputs: trainUsingEM, computeAvgLogLikelihood, forward/backward/align/cluster, printModel,
printAlignment."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

from helpers import IK_CASES, flatten_tables, load_ik, make_model

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.mark.parametrize('case', IK_CASES)
def test_class_api_matches_reference(case, tmp_path):
    g = load_ik(case)
    tmp = str(tmp_path)
    m = make_model(tmp, g)
    assert m.audioFeatDim == g['P'] and len(m.vCorpus) == len(g['feats_list'])
    assert m.aCorpus[0].shape == (len(g['phones_list'][0]), g['P'])
    with pytest.raises(AttributeError):
        with contextlib.redirect_stdout(io.StringIO()):
            m.initializeModel()
            m.printAlignment(os.path.join(tmp, 'early'))      # no conceptCountsA yet (reference :628)
    key = 'W' if g['kind'] == 'linear' else 'mus'
    for it in range(g['n_iter']):
        with contextlib.redirect_stdout(io.StringIO()) as out:
            m.trainUsingEM(1, warmStart=True, printStatus=True)
        assert 'Epoch 0 Average Log Likelihood:' in out.getvalue()
        ll = np.load(os.path.join(tmp, 'm_likelihoods.npy'))[0]
        np.testing.assert_allclose(ll, g['avg_ll'][it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(g['lens'], m.init), g['init_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(g['lens'], m.trans), g['trans_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(m.obs, g['obs_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(getattr(m, key), g['param_%d' % it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(np.concatenate(m.conceptCounts), g['cC_%d' % it], rtol=RTOL, atol=1e-300)
        np.testing.assert_allclose(np.concatenate(m.conceptCountsA), g['cA_%d' % it], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(m.computeAvgLogLikelihood(), float(g['final_ll']), rtol=RTOL)
    # files
    with contextlib.redirect_stdout(io.StringIO()):
        m.printAlignment(os.path.join(tmp, 'ali'))
        m.printModel(os.path.join(tmp, 'model'))
    with open(os.path.join(tmp, 'ali.json')) as f:
        ali = json.load(f)
    keys = {'index', 'image_concepts', 'concept_alignment', 'alignment', 'align_probs', 'is_phoneme'}
    if g['kind'] == 'gaussian':
        keys.add('concept_probs')
    assert set(ali[0].keys()) == keys
    assert np.array_equal(np.concatenate([a['alignment'] for a in ali]), g['alignment'])
    assert np.array_equal(np.concatenate([a['image_concepts'] for a in ali]), g['image_concepts'])
    assert np.array_equal(np.concatenate([a['concept_alignment'] for a in ali]), g['concept_alignment'])
    np.testing.assert_allclose(np.concatenate([np.array(a['align_probs']).ravel() for a in ali]),
                               g['align_probs'], rtol=1e-8)
    txt = open(os.path.join(tmp, 'ali.txt')).read().split('\n\n')
    assert txt[0] == ''.join('%d ' % a for a in ali[0]['alignment'])
    for suffix in ('_initialprobs.txt', '_transitionprobs.txt', '_observationprobs.npy', '_phone2idx.json'):
        assert os.path.exists(os.path.join(tmp, 'model' + suffix))
    line = open(os.path.join(tmp, 'model_initialprobs.txt')).readline()
    m0 = int(g['lens'][0])
    assert line == '%d\t%d\t%f\n' % (m0, 0, m.init[m0][0])
    # single-pair methods
    v0, a0 = m.vCorpus[0], m.aCorpus[0]
    np.testing.assert_allclose(m.forward(v0, a0), g['fwd0'], rtol=RTOL)
    np.testing.assert_allclose(m.backward(v0, a0), g['bwd0'], rtol=RTOL)
    path, probs = m.align(a0, v0)
    T0 = len(g['phones_list'][0])
    assert path == g['alignment'][:T0].tolist()
    cl, scores = m.cluster(a0, v0, path)
    assert cl == g['image_concepts'][:v0.shape[0]].tolist()
    assert np.asarray(scores).shape == (v0.shape[0], g['K'])


def test_train_multi_epoch_and_lr_decay(tmp_path):
    """trainUsingEM(n) in one call == n warm-started single epochs; lr /= 10 after epoch 10 (:260)."""
    g = load_ik('short_toeplitz_linear')
    m1 = make_model(str(tmp_path / 'a'), g) if (tmp_path / 'a').mkdir() is None else None
    m2 = make_model(str(tmp_path / 'b'), g) if (tmp_path / 'b').mkdir() is None else None
    with contextlib.redirect_stdout(io.StringIO()):
        m1.trainUsingEM(3, printStatus=False)
        m2.initializeModel()
        for _ in range(3):
            m2.trainUsingEM(1, warmStart=True, printStatus=False)
    np.testing.assert_allclose(m1.W, m2.W, rtol=1e-12)
    np.testing.assert_allclose(m1.obs, m2.obs, rtol=1e-12)
    np.testing.assert_allclose(m1.W, g['param_2'], rtol=1e-8, atol=1e-12)
    lr0 = m1.lr
    with contextlib.redirect_stdout(io.StringIO()):
        m1.trainUsingEM(10, warmStart=True, printStatus=False)
    assert m1.lr == pytest.approx(lr0 / 10)


@pytest.mark.parametrize('case', ['mixed_twolayer', 'short_twolayer'])
def test_twolayer_class_matches_reference(case, tmp_path):
    """ImagePhoneHMMDNNWordDiscoverer (run_image2phone.py --model_type two-layer), SURVEY 8(f1)."""
    g = load_ik(case)
    tmp = str(tmp_path)
    m = make_model(tmp, g)
    with contextlib.redirect_stdout(io.StringIO()):
        m.initializeModel()
    for it in range(g['n_iter']):
        with contextlib.redirect_stdout(io.StringIO()):
            m.trainUsingEM(1, warmStart=True, printStatus=True)
        ll = np.load(os.path.join(tmp, 'm_likelihoods.npy'))[0]
        np.testing.assert_allclose(ll, g['avg_ll'][it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(g['lens'], m.init), g['init_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(g['lens'], m.trans), g['trans_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(m.obs, g['obs_%d' % it], rtol=RTOL)
        np.testing.assert_allclose(m.W, g['param_%d' % it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(m.V, g['hidden_%d' % it], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(m.computeAvgLogLikelihood(), float(g['final_ll']), rtol=RTOL)
    with contextlib.redirect_stdout(io.StringIO()):
        m.printAlignment(os.path.join(tmp, 'ali'))
        m.printModel(os.path.join(tmp, 'model'))
    ali = json.load(open(os.path.join(tmp, 'ali.json')))
    assert set(ali[0].keys()) == {'index', 'image_concepts', 'alignment', 'cluster_probs', 'align_probs', 'is_phoneme'}
    assert np.array_equal(np.concatenate([a['alignment'] for a in ali]), g['alignment'])
    assert np.array_equal(np.concatenate([a['image_concepts'] for a in ali]), g['image_concepts'])
    np.testing.assert_allclose(np.concatenate([np.array(a['align_probs']).ravel() for a in ali]),
                               g['align_probs'], rtol=1e-8)
    np.testing.assert_allclose(np.concatenate([np.array(a['cluster_probs']).ravel() for a in ali]),
                               g['cluster_probs'], rtol=1e-8, atol=1e-300)
    assert os.path.exists(os.path.join(tmp, 'ali_clusters.txt'))
    assert os.path.exists(os.path.join(tmp, 'model_hiddenweights.npy'))
    v0, a0 = m.vCorpus[0], m.aCorpus[0]
    np.testing.assert_allclose(m.forward(v0, a0), g['fwd0'], rtol=RTOL)
    np.testing.assert_allclose(m.backward(v0, a0), g['bwd0'], rtol=RTOL)
    h0 = m.hiddenLayer(v0)
    assert h0.shape == (v0.shape[0], m.hiddenDim) and np.all(h0 >= 0)
    np.testing.assert_allclose(m.softmaxLayer(h0).sum(1), 1.0, rtol=1e-12)


def test_default_feature_dtype_keeps_float64_inputs_exact(tmp_path):
    """Default config (no feature_dtype key), normalize_vfeat=True: the normalised features are not float32
    values, so the class must keep them in float64 -- tables to 1e-9 and bit-exact Viterbi paths against the
    oracle run on the SAME float64 features (reference :54-55 normalises in float64)."""
    from oracle import image_phone_hmm as orc
    from helpers import write_ik_files
    from multimodalworddiscovery_b200.corpus import resolve_feature_dtype
    from multimodalworddiscovery_b200.hmm_dnn.image_phone_hmm_word_discoverer import ImagePhoneHMMWordDiscoverer
    g = load_ik('mixed_linear')
    tmp = str(tmp_path)
    caps, feats, cfg, extra = write_ik_files(tmp, g)
    cfg.pop('feature_dtype')
    cfg['normalize_vfeat'] = True
    with contextlib.redirect_stdout(io.StringIO()):
        m = ImagePhoneHMMWordDiscoverer(caps, feats, cfg, obsProbFile=extra.get('obs'), modelName=os.path.join(tmp, 'm'))
        m.trainUsingEM(2, printStatus=True)
        m.printAlignment(os.path.join(tmp, 'ali'))
    assert m._eng.feat_is_f64 == 1
    assert resolve_feature_dtype('auto', g['feats_list']) == np.float32      # the un-normalised fixtures ARE float32 values
    vn = [(v.T / np.linalg.norm(v, ord=2, axis=-1)).T for v in g['feats_list']]
    p = orc.initial_params(vn, g['K'], g['P'], 'linear', W=g['param0'], lr=g['lr'], momentum=g['momentum'], obs=g.get('obs0'))
    for _ in range(2):
        p, info = orc.em_iteration(vn, g['phones_list'], p, 'linear')
    np.testing.assert_allclose(m.obs, p['obs'], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(m.W, p['W'], rtol=1e-8, atol=1e-12)
    ali = json.load(open(os.path.join(tmp, 'ali.json')))
    for ex, (v, x) in enumerate(zip(vn, g['phones_list'])):
        n = v.shape[0]
        path, _ = orc.align(orc.posterior_linear(v, p['W']), x, p['obs'], p['init'][n], p['trans'][n])
        assert ali[ex]['alignment'] == path
