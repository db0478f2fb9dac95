"""CPU tests of the host-side logic: packing, bucketing, sharding, table layout, the C-ABI
library's exported symbols, and loud failure without a GPU."""
import contextlib
import io
import os
import re

import numpy as np
import pytest

from helpers import load_ik, make_model
from multimodalworddiscovery_b200 import _lib
from multimodalworddiscovery_b200.corpus import (dense_to_tables, pack_pairs, pack_sorted_arrays,
                                                 shard_positions, tables_to_dense)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _corpus(rng, N=37, D=4, P=6):
    feats = [rng.standard_normal((int(rng.integers(1, 8)), D)) for _ in range(N)]
    phones = [rng.integers(0, P, int(rng.integers(1, 20))) for _ in range(N)]
    return feats, phones


def test_pack_sorted_buckets_and_roundtrip():
    rng = np.random.default_rng(0)
    feats, phones = _corpus(rng)
    pk = pack_pairs(feats, phones, feat_dtype=np.float64)
    n = np.diff(pk.region_off)
    T = np.diff(pk.phone_off)
    assert sorted(pk.order.tolist()) == list(range(len(feats)))
    assert np.all(np.diff(n) >= 0)
    for b in range(len(pk.bucket_n)):
        lo, hi = pk.bucket_lo[b], pk.bucket_lo[b + 1]
        assert np.all(n[lo:hi] == pk.bucket_n[b])
        assert np.all(np.diff(T[lo:hi]) >= 0) and T[lo:hi].max() == pk.bucket_tmax[b]
    for s, ex in enumerate(pk.order):
        np.testing.assert_array_equal(pk.feats[pk.region_off[s]:pk.region_off[s + 1]], feats[ex])
        np.testing.assert_array_equal(pk.phones[pk.phone_off[s]:pk.phone_off[s + 1]], phones[ex])
    assert pk.lens == sorted(set(f.shape[0] for f in feats))
    pk2 = pack_sorted_arrays(pk.region_off, pk.phone_off, pk.feats, pk.phones)
    np.testing.assert_array_equal(pk2.bucket_lo, pk.bucket_lo)
    np.testing.assert_array_equal(pk2.bucket_n, pk.bucket_n)
    np.testing.assert_array_equal(pk2.bucket_tmax, pk.bucket_tmax)


@pytest.mark.parametrize('world', [2, 3, 8])
def test_sharding_partitions_and_balances(world):
    rng = np.random.default_rng(1)
    feats, phones = _corpus(rng, N=203)
    seen, costs = [], []
    for r in range(world):
        pk = pack_pairs(feats, phones, rank=r, world=world)
        seen += pk.order.tolist()
        n = np.diff(pk.region_off).astype(float)
        T = np.diff(pk.phone_off).astype(float)
        costs.append(float(np.sum(T * n ** 3)))
        assert pk.n_pairs_global == 203
        assert pk.lens == sorted(set(f.shape[0] for f in feats))      # global, not per shard
    assert sorted(seen) == list(range(203))
    assert max(costs) / min(costs) < 1.35
    assert shard_positions(10, 1, 4).tolist() == [1, 5, 9]


def test_pack_rejects_bad_pairs():
    rng = np.random.default_rng(2)
    feats, phones = _corpus(rng, N=5)
    with pytest.raises(IndexError):
        pack_pairs(feats, phones[:4] + [np.array([], dtype=int)])
    with pytest.raises(ValueError):
        pack_pairs(feats[:4] + [rng.standard_normal((_lib.NMAX + 1, 4))], phones)
    with pytest.raises(ValueError):
        pack_pairs(feats, phones[:4])


def test_table_layout_roundtrip():
    rng = np.random.default_rng(3)
    init = {m: rng.random(m) for m in (1, 3, 7, 16)}
    trans = {m: rng.random((m, m)) for m in (1, 3, 7, 16)}
    it, tt = tables_to_dense(init, trans)
    assert it.shape == (_lib.NMAX + 1, _lib.NMAX) and tt.shape == (_lib.NMAX + 1, _lib.NMAX ** 2)
    assert tt[3, 1 * 3 + 2] == trans[3][1, 2]
    i2, t2 = dense_to_tables(it, tt, [1, 3, 7, 16])
    for m in init:
        np.testing.assert_array_equal(i2[m], init[m])
        np.testing.assert_array_equal(t2[m], trans[m])


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every function include/mwd_b200.h declares."""
    hdr = open(os.path.join(ROOT, 'include', 'mwd_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(mwd_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.mwd_version() >= 100
    assert lib.mwd_ik_counts_len(65, 49) == 65 * 49 + 17 * 16 + 17 * 256 + 1


def test_struct_layout_matches_header():
    import ctypes as C
    lib = _lib.load()
    for which, st in enumerate((_lib.Geometry, _lib.IkProblem, _lib.PartialSizes, _lib.IkMstepArgs,
                                _lib.HmmProblem, _lib.HmmMstepArgs)):
        assert lib.mwd_abi_sizeof(which) == C.sizeof(st)


def test_class_constructs_on_cpu_and_compute_fails_loudly(tmp_path):
    """Host-side reading/initialisation works anywhere; kernels never silently fall back."""
    import torch
    g = load_ik('tiny_linear')
    m = make_model(str(tmp_path), g)
    with contextlib.redirect_stdout(io.StringIO()):
        m.initializeModel()
    assert sorted(m.lenProb) == [2] and m.obs.shape == (3, 3) and m.W.shape == (3, 4)
    np.testing.assert_array_equal(m.aCorpus[1], np.array([[0, 1., 0], [0, 0, 1.]]))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.MwdError):
            m.computeAvgLogLikelihood()


def test_pack_audio_pairs_shards_cover_the_corpus():
    """engine_audio.pack_audio_pairs: frames follow the (n, T)-sorted order of their pairs, the identity
    phone sequence indexes them, and the round-robin shards of two ranks partition the corpus."""
    from multimodalworddiscovery_b200.engine_audio import pack_audio_pairs
    rng = np.random.default_rng(0)
    feats = [rng.standard_normal((int(rng.integers(1, 6)), 4)) for _ in range(11)]
    audio = [rng.standard_normal((int(rng.integers(1, 9)), 3)) for _ in range(11)]
    seen = []
    for rank in range(2):
        pk, aud = pack_audio_pairs(feats, audio, feat_dtype=np.float64, rank=rank, world=2)
        assert np.array_equal(pk.phones, np.arange(pk.n_phones_total))
        for s, ex in enumerate(pk.order):
            np.testing.assert_array_equal(aud[pk.phone_off[s]:pk.phone_off[s + 1]], audio[int(ex)])
            np.testing.assert_array_equal(pk.feats[pk.region_off[s]:pk.region_off[s + 1]], feats[int(ex)])
        n = np.diff(pk.region_off)
        assert np.all(np.diff(n) >= 0)
        seen += [int(e) for e in pk.order]
    assert sorted(seen) == list(range(11))


def test_segembed_batched_embedding_equals_per_segment_oracle():
    """SegEmbedHMMWordDiscoverer.getSentEmbeds resamples all equal-length segments of an utterance in one
    batched FFT call; the values must be those of the per-segment restatement (oracle/segembed_hmm.py,
    reference hmm/audio_segembed_hmm_word_discoverer.py:114-155)."""
    from oracle import segembed_hmm as sh
    from multimodalworddiscovery_b200.hmm.audio_segembed_hmm_word_discoverer import (SegEmbedHMMWordDiscoverer,
                                                                                      _clip_landmarks)
    rng = np.random.default_rng(2)
    m = SegEmbedHMMWordDiscoverer.__new__(SegEmbedHMMWordDiscoverer)
    m.embedDim, m.frameDim, m.featDim = 120, 12, 13
    x = rng.standard_normal((300, 13))
    seg = [0, 7, 19, 26, 26 + 12, 60, 67, 100, 131, 138]          # repeated lengths 7 and 12
    mine = m.getSentEmbeds(x, seg, frameDim=12)
    ref = sh.sent_embeds(x, seg, 120, 12)
    assert mine.shape == (len(seg) - 1, 120)
    np.testing.assert_allclose(mine, ref, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(m.embed(x[7:19], frameDim=12), ref[1], rtol=1e-6, atol=1e-6)
    assert m.getSentDurations(seg) == list(np.diff(seg))
    # landmarks past maxLen collapse into one closing boundary (reference :68-76)
    assert _clip_landmarks(np.array([0, 5, 9, 14, 30]), 10) == [0, 5, 9, 10]
    assert _clip_landmarks(np.array([0, 5, 9]), 10) == [0, 5, 9]


def test_radix_digit_argsort_matches_numpy_stable():
    """The postings index is built with LSD passes over 16-bit digits; it must be THE stable order."""
    from multimodalworddiscovery_b200.engine_hmm import _stable_argsort_u16_digits
    rng = np.random.default_rng(3)
    for bound in (1, 7, 65536, 65537, 1546 * 69, (1 << 33) + 11):
        keys = rng.integers(0, bound, 50000).astype(np.int64)
        assert np.array_equal(_stable_argsort_u16_digits(keys, bound), np.argsort(keys, kind='stable'))
    assert _stable_argsort_u16_digits(np.zeros(0, np.int64), 5).shape == (0,)


def test_mixed_precision_spec_parsing():
    """modelConfigs['posterior_precision'] -> MWD_MIXED_* bits: 'mixed' is everything that passes the 1e-5 gate
    ('all' is kept as a synonym), subsets are '+'-joined, the default is the reference arithmetic."""
    every = _lib.MIXED_CONCEPT | _lib.MIXED_POSTERIOR | _lib.MIXED_GRAD | _lib.MIXED_RECURSION
    for spec in (None, False, 0, 'float64'):
        assert _lib.mixed_bits(spec) == 0
    assert _lib.mixed_bits('mixed') == _lib.mixed_bits('all') == _lib.mixed_bits(True) == every
    assert _lib.mixed_bits('posterior+grad') == _lib.MIXED_POSTERIOR | _lib.MIXED_GRAD
    assert _lib.mixed_bits('recursion + concept') == _lib.MIXED_RECURSION | _lib.MIXED_CONCEPT
    assert _lib.mixed_bits(every | 64) == every
    with pytest.raises(KeyError):
        _lib.mixed_bits('tensor')


def test_library_path_override(tmp_path):
    """MWD_B200_LIB names another build of the library (A/B timing); a wrong path fails loudly, there is no fallback."""
    import subprocess
    import sys
    bogus = str(tmp_path / 'libmwd_other.so')
    code = ('from multimodalworddiscovery_b200 import _lib\n'
            'assert _lib.LIB_PATH == %r, _lib.LIB_PATH\n'
            'try:\n'
            '    _lib.load()\n'
            'except _lib.MwdError as e:\n'
            '    print("LOUD", e)\n' % bogus)
    env = dict(os.environ, MWD_B200_LIB=bogus, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert 'LOUD' in out.stdout and 'no CPU fallback' in out.stdout
