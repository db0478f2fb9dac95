"""oracle/segembed_hmm.py (config 4) against vectors generated from the pieces the reference ships
(tests/golden/make_golden_seg.py): gaussian()/gmmProb() of smt/audio_gmm_word_discoverer.py:53-106, the mean
update of GMMWordDiscoverer.updateTranslationDensities :339-375 and embed()/getSentEmbeds() of
hmm/audio_segembed_hmm_word_discoverer.py:114-155.  CPU only."""
import os

import numpy as np

from helpers import GOLDEN
from oracle import segembed_hmm as sh


def _g():
    z = np.load(os.path.join(GOLDEN, 'seg_pieces.npz'))
    return {k: z[k] for k in z.files}


def test_log_gauss_equals_reference_gaussian():
    g = _g()
    for m in range(g['g_means'].shape[0]):
        np.testing.assert_allclose(sh.log_gauss(g['g_x'], g['g_means'][m], g['g_var'][m]), g['g_loggauss'][m],
                                   rtol=1e-13, atol=1e-11)


def test_mixture_emission_equals_reference_gmmprob():
    g = _g()
    M = g['g_means'].shape[0]
    # one word (id 0) with M mixtures, one state
    lb, resp = sh.emission(g['g_x'], [0], g['g_lprior'][None], g['g_means'][None], g['g_var'][None])
    np.testing.assert_allclose(lb[:, 0], g['g_gmm'], rtol=1e-13, atol=1e-11)
    np.testing.assert_allclose(lb[0, 0], float(g['g_gmm_row0']), rtol=1e-13)
    np.testing.assert_allclose(np.exp(resp[:, 0]).sum(1), 1.0, rtol=1e-12)
    lb2, _ = sh.emission(g['g_x'], [0], g['g_lprior'][None], g['g_means'][None], 0.02 * np.ones((1, M, g['g_x'].shape[1])))
    np.testing.assert_allclose(lb2[:, 0], g['g_gmm_fixedvar'], rtol=1e-13, atol=1e-9)


def test_mean_update_equals_reference_mstep():
    g = _g()
    Vt, M, D = g['m_means'].shape
    to, xo = g['m_tgt_off'], g['m_x_off']
    embs, tgt, logw = [], [], []
    pos = 0
    for u in range(int(g['m_n_utts'])):
        e = g['m_tgt'][to[u]:to[u + 1]]
        x = g['m_x'][xo[u]:xo[u + 1]]
        T = x.shape[0]
        lw = np.empty((T, len(e), M))
        for j in range(len(e)):
            lw[:, j, :] = g['m_logw'][pos:pos + M * T].reshape(M, T).T      # stored (mixture, t)
            pos += M * T
        embs.append(x)
        tgt.append(e)
        logw.append(lw)
    assert pos == len(g['m_logw'])
    w_sum, x_sum, _ = sh.weighted_stats(embs, tgt, logw, Vt, M)
    means, _, ok = sh.means_from_stats(np.zeros((Vt, M, D)), w_sum, x_sum)
    assert ok.all()
    np.testing.assert_allclose(means, g['m_means'], rtol=1e-11, atol=1e-13)


def test_embed_equals_reference_embed():
    g = _g()
    seg = g['e_seg'].tolist()
    np.testing.assert_allclose(sh.sent_embeds(g['e_utt'], seg, 120, 12), g['e_embeds'], rtol=0, atol=0)
    np.testing.assert_allclose(sh.embed(g['e_utt'][0:7], 120, 12), g['e_embed_first'], rtol=0, atol=0)
    np.testing.assert_allclose(sh.sent_embeds(g['e_utt'], seg, 560, 12), g['e_embeds_560'], rtol=0, atol=0)


def test_class_batched_embedding_equals_reference_embed():
    """The product class resamples all equal-length segments of an utterance in one FFT call; the values
    must be those of the reference's per-segment calls (host-side code, no GPU needed)."""
    from multimodalworddiscovery_b200.hmm.audio_segembed_hmm_word_discoverer import SegEmbedHMMWordDiscoverer
    g = _g()
    obj = object.__new__(SegEmbedHMMWordDiscoverer)
    obj.embedDim, obj.frameDim, obj.featDim = 120, 12, 14
    got = obj.getSentEmbeds(g['e_utt'], g['e_seg'].tolist(), frameDim=12)
    np.testing.assert_allclose(got, g['e_embeds'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(obj.embed(g['e_utt'][0:7], frameDim=12), g['e_embed_first'], rtol=1e-6, atol=1e-6)
