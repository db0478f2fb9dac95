"""Segment-embedding HMM (config 4): CUDA vs oracle/segembed_hmm.py.  The reference class cannot run end to
end as shipped, so the oracle is pinned piece by piece to the reference code that exists -- gaussian() /
gmmProb(), the GMM mean update, embed() (tests/test_oracle_seg_golden.py, tests/golden/seg_pieces.npz) and
the log-domain recursion (hmm_*_log.npz) -- and the CUDA emission is also checked against those vectors
directly (last test of this file)."""
import numpy as np
import pytest

from helpers import flatten_tables
from oracle import plain_hmm as ph
from oracle import segembed_hmm as sh

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.mark.parametrize('M,D,update_var', [(1, 12, False), (2, 24, False), (3, 120, True)])
def test_segment_hmm_matches_oracle(M, D, update_var):
    from multimodalworddiscovery_b200.engine_hmm import SegmentHMMEngine
    rng = np.random.default_rng(7 + M)
    Vt, N = 9, 30
    cent = 2.0 * rng.standard_normal((Vt, D))
    embs, tgt = [], []
    for _ in range(N):
        n = int(rng.integers(1, 6))
        e = np.concatenate([[0], rng.integers(1, Vt, n)])
        S = int(rng.integers(2, 14))
        st = rng.integers(0, len(e), S)
        embs.append((cent[e[st]] + 0.4 * rng.standard_normal((S, D))).astype(np.float32).astype(np.float64))
        tgt.append(e)
    lens = sorted({len(e) for e in tgt})
    p = dict(init={m: np.log(1. / m) * np.ones(m) for m in lens},
             trans={m: np.log(1. / m) * np.ones((m, m)) for m in lens},
             lprior=np.log(np.ones((Vt, M)) / M), means=cent[:, None, :] + 0.5 * rng.standard_normal((Vt, M, D)),
             var=0.3 * np.ones((Vt, M, D)))
    acc = ph.LogAccumulators(lens, Vt, 1)
    eng = SegmentHMMEngine(tgt, embs, Vt, M, emb_dtype=np.float64)
    eng.set_chain_params(p['init'], p['trans'])
    eng.set_emission_params(p['lprior'], p['means'], p['var'])
    for it in range(3):
        p, info = sh.em_iteration(embs, tgt, p, acc, update_var=update_var)
        eng.em_iteration(update_var=update_var)
        ll = float(eng.loglik_sum()) / N
        np.testing.assert_allclose(ll, info['avg_ll'], rtol=RTOL)
        init, trans, lprior, means, var = eng.get_all_params()
        np.testing.assert_allclose(flatten_tables(lens, init), flatten_tables(lens, p['init']), rtol=RTOL)
        np.testing.assert_allclose(flatten_tables(lens, trans), flatten_tables(lens, p['trans']), rtol=RTOL)
        np.testing.assert_allclose(means, p['means'], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(lprior, p['lprior'], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(var, p['var'], rtol=1e-7, atol=1e-10)
    ali, ap = eng.align()
    ali = ali.cpu().numpy()
    pk = eng.pk
    for s, ex in enumerate(pk.order):
        path, probs = sh.align(embs[ex], tgt[ex], p)
        mine = ali[pk.src_off[s]:pk.src_off[s + 1]].tolist()
        if len(set(tgt[ex].tolist())) == len(tgt[ex]):
            assert mine == path                      # distinct concepts: no symmetric ties
        else:
            # a concept repeated in the caption gives two states with identical emissions; both
            # paths must then have the same Viterbi score
            e = tgt[ex]
            lb, _ = sh.emission(embs[ex], e, p['lprior'], p['means'], p['var'])
            lpi, lA = p['init'][len(e)], p['trans'][len(e)]

            def score(q):
                v = lpi[q[0]] + lb[0, q[0]]
                for t in range(1, len(q)):
                    v += lA[q[t - 1], q[t]] + lb[t, q[t]]
                return v
            assert score(mine) == pytest.approx(score(path), rel=1e-9)


def test_segembed_wrapper_end_to_end(tmp_path):
    """SegEmbedHMMWordDiscoverer(AudioHMMWordDiscoverer, ...) as run_audio.py:185-194 builds it:
    synthetic MFCC utterances + landmarks -> embeddings -> EM -> printAlignment files."""
    import contextlib
    import io
    import json
    from multimodalworddiscovery_b200.hmm.audio_hmm_word_discoverer import AudioHMMWordDiscoverer
    from multimodalworddiscovery_b200.hmm.audio_segembed_hmm_word_discoverer import SegEmbedHMMWordDiscoverer
    rng = np.random.default_rng(11)
    words = ['dog', 'ball', 'tree', 'car']
    protos = {w: rng.standard_normal((10, 12)) for w in words + ['NULL']}
    feats, lms, caps = {}, {}, []
    for u in range(12):
        concepts = list(rng.choice(words, size=int(rng.integers(1, 4)), replace=False))
        seq = [rng.choice(['NULL'] + concepts) for _ in range(int(rng.integers(3, 8)))]
        frames, bounds = [], [0]
        for w in seq:
            L = int(rng.integers(6, 15))
            idx = np.linspace(0, 9, L).astype(int)
            frames.append(protos[w][idx] + 0.1 * rng.standard_normal((L, 12)))
            bounds.append(bounds[-1] + L)
        feats['arr_%d' % u] = np.concatenate(frames)
        lms['arr_%d' % u] = np.array(bounds)
        caps.append(' '.join(concepts))
    np.savez(str(tmp_path / 'src.npz'), **feats)
    np.savez(str(tmp_path / 'lm.npz'), **lms)
    with open(str(tmp_path / 'trg.txt'), 'w') as f:
        f.write('\n'.join(caps))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m = SegEmbedHMMWordDiscoverer(AudioHMMWordDiscoverer, 1, 12, 120, str(tmp_path / 'src.npz'),
                                      str(tmp_path / 'trg.txt'), landmarkFile=str(tmp_path / 'lm.npz'))
        assert m.embeddings[0].shape[1] == 120
        m.trainUsingEM(4, writeModel=False)
        m.printAlignment(str(tmp_path / 'ali'))
    lls = [float(ln.split(':')[-1]) for ln in buf.getvalue().split('\n') if 'Average Log Likelihood' in ln]
    # (the reference's count quirks -- last-t-only transition counts, un-normalised init counts --
    # do not make this EM monotone, so only finiteness is asserted)
    assert len(lls) == 4 and np.all(np.isfinite(lls))
    ali = json.load(open(str(tmp_path / 'ali.json')))
    assert len(ali) == 12 and ali[0]['is_audio'] and ali[0]['image_concepts'][0] == 'NULL'
    assert len(ali[3]['alignment']) == int(lms['arr_3'][-1])          # one state index per frame


def test_cuda_gaussian_emission_equals_reference_gmmprob():
    """mwd_hmm_gauss_emission against vectors produced by the reference's own gaussian()/gmmProb()
    (smt/audio_gmm_word_discoverer.py:53-106; tests/golden/seg_pieces.npz)."""
    import os
    from helpers import GOLDEN
    from multimodalworddiscovery_b200.engine_hmm import SegmentHMMEngine
    z = np.load(os.path.join(GOLDEN, 'seg_pieces.npz'))
    x, means, var, lprior = z['g_x'], z['g_means'], z['g_var'], z['g_lprior']
    M = means.shape[0]
    eng = SegmentHMMEngine([np.array([0])], [x], 1, M, emb_dtype=np.float64)
    eng.set_chain_params({1: np.zeros(1)}, {1: np.zeros((1, 1))})
    eng.set_emission_params(lprior[None], means[None], var[None])
    eng.emission()
    np.testing.assert_allclose(eng.emis.cpu().numpy()[:x.shape[0]], z['g_gmm'], rtol=1e-12, atol=1e-10)
    eng.set_emission_params(lprior[None], means[None], 0.02 * np.ones((1, M, x.shape[1])))
    eng.emission()
    np.testing.assert_allclose(eng.emis.cpu().numpy()[:x.shape[0]], z['g_gmm_fixedvar'], rtol=1e-12, atol=1e-8)
