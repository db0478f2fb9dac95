"""Host-side packing of caption-image pairs for the (region i, concept k)-state kernels.

The reference keeps ``vCorpus`` (list of (n, D) arrays) and ``aCorpus`` (list of one-hot (T, P)
arrays) and loops over them in Python (hmm_dnn/image_phone_hmm_word_discoverer.py:224).  Here the
pairs are sorted by (n, T) into length buckets, stored CSR-style (see include/mwd_b200.h) and, for
multi-GPU runs, dealt round-robin over the sorted order so that every rank gets the same mix of
(n, T) -- i.e. equal sum of T*K*n^3 work, not merely equal pair counts.

Pure NumPy; no device code here.
"""
import numpy as np

from ._lib import NMAX


class PackedPairs(object):
    """One rank's shard of the corpus in kernel layout.

    Attributes
      order        (N,) int64  original corpus index of the pair stored at sorted position s
      region_off   (N+1,) int32, phone_off (N+1,) int32
      feats        (R, D) float32|float64, phones (Ttot,) int32
      bucket_n     (nb,) int32, bucket_lo (nb+1,) int64, bucket_tmax (nb,) int32
      lens         sorted distinct n over the WHOLE corpus (all ranks) -- reference ``lenProb`` keys
      n_pairs_global
    """

    def __init__(self):
        self.order = None

    @property
    def n_pairs(self):
        return len(self.order)

    @property
    def n_regions(self):
        return int(self.region_off[-1])

    @property
    def n_phones_total(self):
        return int(self.phone_off[-1])

    @property
    def t_max(self):
        return int(self.bucket_tmax.max()) if len(self.bucket_tmax) else 0

    def ap_offsets(self):
        """Offsets of each pair's (T, n) alignProbs block in the ragged align_probs buffer."""
        n = np.diff(self.region_off).astype(np.int64)
        T = np.diff(self.phone_off).astype(np.int64)
        return np.concatenate([[0], np.cumsum(n * T)]).astype(np.int64)


def shard_positions(n_sorted, rank, world):
    """Sorted positions owned by ``rank``: round-robin over the (n, T)-sorted order."""
    return np.arange(rank, n_sorted, world, dtype=np.int64)


def pack_pairs(feats_list, phones_list, feat_dtype=np.float32, rank=0, world=1):
    """Sort by (n, T), shard, and lay out one rank's pairs.

    feats_list[ex]: (n, D) array; phones_list[ex]: (T,) integer phone ids.
    """
    N = len(feats_list)
    if N != len(phones_list):
        raise ValueError('corpus mismatch: %d images vs %d captions' % (N, len(phones_list)))
    if N == 0:
        raise ValueError('empty corpus')
    ns = np.array([f.shape[0] for f in feats_list], dtype=np.int64)
    Ts = np.array([len(x) for x in phones_list], dtype=np.int64)
    if ns.min() < 1:
        raise ValueError('pair %d has no image regions' % int(np.argmin(ns)))
    if ns.max() > NMAX:
        raise ValueError('pair %d has %d regions; this build supports at most %d'
                         % (int(np.argmax(ns)), int(ns.max()), NMAX))
    if Ts.min() < 1:
        # the reference raises IndexError in forward() on an empty caption (:287)
        raise IndexError('pair %d has an empty caption' % int(np.argmin(Ts)))
    D = feats_list[0].shape[1]
    # stable sort by (n, T): np.lexsort sorts by last key first
    sorted_idx = np.lexsort((np.arange(N), Ts, ns))
    mine = sorted_idx[shard_positions(N, rank, world)]

    pk = PackedPairs()
    pk.order = mine.astype(np.int64)
    pk.lens = sorted(int(v) for v in np.unique(ns))
    pk.n_pairs_global = N
    n_m, T_m = ns[mine], Ts[mine]
    pk.region_off = np.concatenate([[0], np.cumsum(n_m)]).astype(np.int32)
    pk.phone_off = np.concatenate([[0], np.cumsum(T_m)]).astype(np.int32)
    if int(np.sum(T_m)) >= 2 ** 31 or int(np.sum(n_m)) >= 2 ** 31:
        raise ValueError('shard too large for int32 offsets; use more ranks')
    if len(mine):
        pk.feats = np.ascontiguousarray(
            np.concatenate([np.asarray(feats_list[i]).reshape(-1, D) for i in mine], axis=0),
            dtype=feat_dtype)
        pk.phones = np.concatenate([np.asarray(phones_list[i]) for i in mine]).astype(np.int32)
    else:
        pk.feats = np.zeros((0, D), dtype=feat_dtype)
        pk.phones = np.zeros((0,), dtype=np.int32)
    # buckets of equal n (contiguous because of the sort)
    bn, blo, btm = [], [0], []
    for s in range(len(mine)):
        if s == 0 or n_m[s] != n_m[s - 1]:
            if s:
                blo.append(s)
                btm.append(int(T_m[blo[-2]:s].max()))
            bn.append(int(n_m[s]))
    if len(mine):
        blo.append(len(mine))
        btm.append(int(T_m[blo[-2]:].max()))
    pk.bucket_n = np.array(bn, dtype=np.int32)
    pk.bucket_lo = np.array(blo, dtype=np.int64)
    pk.bucket_tmax = np.array(btm, dtype=np.int32)
    return pk


def pack_sorted_arrays(region_off, phone_off, feats, phones, lens=None, n_pairs_global=None):
    """Wrap arrays that are ALREADY sorted by (n, T) and CSR-packed (bench / streaming path)."""
    pk = PackedPairs()
    N = len(region_off) - 1
    pk.order = np.arange(N, dtype=np.int64)
    pk.region_off = np.asarray(region_off, dtype=np.int32)
    pk.phone_off = np.asarray(phone_off, dtype=np.int32)
    pk.feats = feats
    pk.phones = phones
    n_m = np.diff(pk.region_off)
    T_m = np.diff(pk.phone_off)
    if N and (np.any(np.diff(n_m) < 0)):
        raise ValueError('pairs must be sorted by n')
    change = np.flatnonzero(np.diff(n_m)) + 1
    blo = np.concatenate([[0], change, [N]]).astype(np.int64)
    pk.bucket_lo = blo
    pk.bucket_n = n_m[blo[:-1]].astype(np.int32)
    pk.bucket_tmax = np.array([int(T_m[blo[b]:blo[b + 1]].max()) for b in range(len(blo) - 1)],
                              dtype=np.int32)
    pk.lens = sorted(int(v) for v in np.unique(n_m)) if lens is None else list(lens)
    pk.n_pairs_global = N if n_pairs_global is None else int(n_pairs_global)
    return pk


def tables_to_dense(init, trans):
    """Reference dicts ``init[m] (m,)``, ``trans[m] (m,m)`` -> dense kernel tables."""
    it = np.zeros((NMAX + 1, NMAX), dtype=np.float64)
    tt = np.zeros((NMAX + 1, NMAX * NMAX), dtype=np.float64)
    for m, v in init.items():
        it[int(m), :int(m)] = np.asarray(v, dtype=np.float64)
    for m, v in trans.items():
        m = int(m)
        tt[m, :m * m] = np.asarray(v, dtype=np.float64).reshape(-1)
    return it, tt


def dense_to_tables(it, tt, lens):
    init = {int(m): np.array(it[int(m), :int(m)], dtype=np.float64) for m in lens}
    trans = {int(m): np.array(tt[int(m), :int(m) * int(m)], dtype=np.float64).reshape(int(m), int(m))
             for m in lens}
    return init, trans


def resolve_feature_dtype(spec, *array_lists):
    """Device storage type of the feature matrices.  ``spec``: 'float64' | 'float32' | 'auto'.
    'auto' (the class default) keeps the reference's float64 arithmetic exact: float32 storage is chosen
    only when EVERY value survives the round trip through float32 unchanged (float32 .npz inputs, or
    float64 arrays holding float32-representable values); anything else -- float64 .npz files, the
    output of ``normalize_vfeat`` -- stays float64, so the tables and the bit-exact Viterbi paths are
    those of the reference's inputs, not of rounded ones.  'float32' is an explicit opt-in to rounding."""
    if spec == 'float64':
        return np.float64
    if spec == 'float32':
        return np.float32
    if spec != 'auto':
        raise ValueError("feature_dtype must be 'auto', 'float32' or 'float64', not %r" % (spec,))
    for arrays in array_lists:
        for a in arrays:
            a = np.asarray(a)
            if a.dtype == np.float32 or a.size == 0:
                continue
            if a.dtype.kind != 'f' and a.dtype.kind not in 'iub':
                return np.float64
            with np.errstate(over='ignore', invalid='ignore'):
                if not np.array_equal(a, a.astype(np.float32)):
                    return np.float64
    return np.float32
