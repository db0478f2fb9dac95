"""Device state + EM driver of the plain-state HMM word discoverers (hmm/ classes).

Host side: packing by (n states, T), rank sharding, and the static postings index that turns the
observation-count scatter into a deterministic segmented reduction.  All computation is CUDA
behind include/mwd_b200.h (mwd_hmm_*); PyTorch is device memory / streams / collectives only.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import HmmMstepArgs, HmmProblem, MwdError, NMAX
from .corpus import dense_to_tables, shard_positions, tables_to_dense
from .dist import fixed_order_allreduce


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


def _stable_argsort_u16_digits(keys, bound):
    """Stable argsort of non-negative integer keys < bound, as LSD passes over 16-bit digits: numpy's
    stable sort is a radix sort only for 16-bit types, and the postings index of a 1 M-pair corpus has
    1.5e8 keys (2.8x faster than a merge sort of the int64 keys)."""
    idx = None
    shift = 0
    while shift == 0 or (bound - 1) >> shift:
        digit = ((keys >> shift) & 0xffff).astype(np.uint16)
        if idx is None:
            idx = np.argsort(digit, kind='stable')
        else:
            idx = idx[np.argsort(digit[idx], kind='stable')]
        shift += 16
    return idx.astype(np.int64)


class PackedSentences(object):
    """(concept states, phone tokens) pairs sorted by (n, T), CSR-packed, with postings."""

    def __init__(self, tgt_ids, src_ids, n_src_types, rank=0, world=1):
        N = len(tgt_ids)
        if N == 0 or N != len(src_ids):
            raise ValueError('corpus mismatch: %d target vs %d source sentences' % (N, len(src_ids)))
        ns = np.array([len(e) for e in tgt_ids], dtype=np.int64)
        Ts = np.array([len(f) for f in src_ids], dtype=np.int64)
        if ns.min() < 1 or Ts.min() < 1:
            raise IndexError('empty sentence in pair %d' % int(np.argmin(np.minimum(ns, Ts))))
        if ns.max() > NMAX:
            raise ValueError('%d states in pair %d; this build supports at most %d'
                             % (int(ns.max()), int(np.argmax(ns)), NMAX))
        order = np.lexsort((np.arange(N), Ts, ns))
        mine = order[shard_positions(N, rank, world)]
        self.order = mine.astype(np.int64)
        self.lens = sorted(int(v) for v in np.unique(ns))
        self.n_pairs_global = N
        n_m, T_m = ns[mine], Ts[mine]
        self.tgt_off = np.concatenate([[0], np.cumsum(n_m)]).astype(np.int32)
        self.src_off = np.concatenate([[0], np.cumsum(T_m)]).astype(np.int32)
        self.slot_off = np.concatenate([[0], np.cumsum(n_m * T_m)]).astype(np.int64)
        self.ap_off = np.concatenate([[0], np.cumsum(n_m * (T_m - 1))]).astype(np.int64)
        self.tgt = (np.concatenate([np.asarray(tgt_ids[i]) for i in mine]) if len(mine)
                    else np.zeros(0)).astype(np.int32)
        self.src = (np.concatenate([np.asarray(src_ids[i]) for i in mine]) if len(mine)
                    else np.zeros(0)).astype(np.int32)
        # launch groups: equal state count.  (Splitting them further by caption-length range shrinks the
        # recursion kernels' shared-memory slabs, but the extra serial launches and their tails cost more
        # than the occupancy returns: 2.66 -> 4.42 ms per 200k pairs with six ranges, profiles/r01_bench_history.md)
        key = n_m
        change = np.flatnonzero(np.diff(key)) + 1
        blo = np.concatenate([[0], change, [len(mine)]]).astype(np.int64) if len(mine) else np.zeros(1, np.int64)
        self.bucket_lo = blo
        self.bucket_n = n_m[blo[:-1]].astype(np.int32)
        self.bucket_tmax = np.array([int(T_m[blo[b]:blo[b + 1]].max()) for b in range(len(blo) - 1)],
                                    dtype=np.int32)
        self.Vf = int(n_src_types)

    @property
    def n_pairs(self):
        return len(self.order)

    @property
    def n_slots(self):
        return int(self.slot_off[-1])

    @property
    def t_max(self):
        return int(self.bucket_tmax.max()) if len(self.bucket_tmax) else 0

    def postings(self, n_tgt_types):
        """Slots (pair, t, i) sorted by table entry (concept e_i, phone f_t): (idx, off)."""
        n_p = np.diff(self.tgt_off).astype(np.int64)
        T_p = np.diff(self.src_off).astype(np.int64)
        pair_of_pos = np.repeat(np.arange(self.n_pairs), T_p)          # pair of each phone position
        rep = n_p[pair_of_pos]                                         # slots per phone position
        pos_of_slot = np.repeat(np.arange(len(self.src)), rep)
        first_slot = np.concatenate([[0], np.cumsum(rep)])[:-1]
        i_of_slot = np.arange(self.n_slots) - np.repeat(first_slot, rep)
        tgt_idx = self.tgt_off[pair_of_pos[pos_of_slot]].astype(np.int64) + i_of_slot
        keys = self.tgt[tgt_idx].astype(np.int64) * self.Vf + self.src[pos_of_slot].astype(np.int64)
        self.slot_row = pos_of_slot.astype(np.int32)      # source row (phone position / segment) of each slot
        self.row_pair = pair_of_pos.astype(np.int32)      # pair of each source row
        idx = _stable_argsort_u16_digits(keys, n_tgt_types * self.Vf)
        off = np.searchsorted(keys[idx], np.arange(n_tgt_types * self.Vf + 1)).astype(np.int64)
        return idx, off


class PlainHMMEngine(object):
    def __init__(self, packed, n_tgt_types, n_src_types, log_domain, device=None, process_group=None):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise MwdError('no CUDA device: the mwd_b200 engine has no CPU fallback')
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.pk = packed
        self.Vt, self.Vf = int(n_tgt_types), int(n_src_types)
        self.log = bool(log_domain)
        self.pg = process_group
        dev, f64 = self.device, torch.float64

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

        self.tgt_off, self.tgt = up(packed.tgt_off), up(packed.tgt)
        self.src_off, self.src = up(packed.src_off), up(packed.src)
        self.slot_off, self.ap_off = up(packed.slot_off), up(packed.ap_off)
        idx, off = packed.postings(self.Vt)
        self.post_idx, self.post_off = up(idx), up(off)
        self._bucket_n = np.ascontiguousarray(packed.bucket_n, dtype=np.int32)
        self._bucket_lo = np.ascontiguousarray(packed.bucket_lo, dtype=np.int64)
        self._bucket_tmax = np.ascontiguousarray(packed.bucket_tmax, dtype=np.int32)
        self._lens = np.ascontiguousarray(np.array(packed.lens, dtype=np.int32))
        self.init_t = torch.zeros((NMAX + 1, NMAX), dtype=f64, device=dev)
        self.trans_t = torch.zeros((NMAX + 1, NMAX * NMAX), dtype=f64, device=dev)
        self.obs = torch.full((self.Vt, self.Vf), float('nan'), dtype=f64, device=dev)
        self.pair_ll = torch.zeros((max(packed.n_pairs, 1),), dtype=f64, device=dev)
        self.post = torch.zeros((max(packed.n_slots, 1),), dtype=f64, device=dev)
        self.warps = int(self.lib.mwd_hmm_warps())
        self.part_init = torch.zeros((self.warps, NMAX + 1, NMAX), dtype=f64, device=dev)
        self.part_trans = torch.zeros((self.warps, NMAX + 1, NMAX * NMAX), dtype=f64, device=dev)
        self.counts_len = int(self.lib.mwd_hmm_counts_len(self.Vt, self.Vf))
        self.counts = torch.zeros((self.counts_len,), dtype=f64, device=dev)
        self.acc = None
        self.reset_accumulators()

    def reset_accumulators(self):
        """The log class keeps counts over epochs (reference lists created outside the epoch loop)."""
        if self.log:
            self.acc = self.torch.full((self.counts_len,), float('-inf'), dtype=self.torch.float64,
                                       device=self.device)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _problem(self, alpha=None, beta=None):
        pk = self.pk
        p = HmmProblem()
        p.n_pairs, p.n_slots = pk.n_pairs, pk.n_slots
        p.n_tgt_types, p.n_src_types, p.t_max = self.Vt, self.Vf, pk.t_max
        p.log_domain = 1 if self.log else 0
        p.n_buckets = len(self._bucket_n)
        p.bucket_n, p.bucket_lo = _np_ptr(self._bucket_n), _np_ptr(self._bucket_lo)
        p.bucket_tmax = _np_ptr(self._bucket_tmax)
        p.tgt_off, p.tgt, p.src_off, p.src = _ptr(self.tgt_off), _ptr(self.tgt), _ptr(self.src_off), _ptr(self.src)
        p.slot_off = _ptr(self.slot_off)
        p.init, p.trans, p.obs = _ptr(self.init_t), _ptr(self.trans_t), _ptr(self.obs)
        p.pair_ll, p.post = _ptr(self.pair_ll), _ptr(self.post)
        p.part_init, p.part_trans = _ptr(self.part_init), _ptr(self.part_trans)
        p.alpha_out, p.beta_out = _ptr(alpha), _ptr(beta)
        p.emis = _ptr(getattr(self, 'emis', None))
        p.n_src_rows = len(pk.src)
        p.row_pair, p.slot_row = _ptr(getattr(self, 'row_pair', None)), _ptr(getattr(self, 'slot_row', None))
        return p

    # ------------------------------------------------------------------ parameters
    def set_params(self, init, trans, obs_dense):
        torch = self.torch
        it, tt = tables_to_dense(init, trans)
        self.init_t.copy_(torch.from_numpy(it))
        self.trans_t.copy_(torch.from_numpy(tt))
        self.obs.copy_(torch.from_numpy(np.ascontiguousarray(obs_dense, dtype=np.float64)))

    def get_params(self):
        init, trans = dense_to_tables(self.init_t.cpu().numpy(), self.trans_t.cpu().numpy(), self.pk.lens)
        return init, trans, self.obs.cpu().numpy().copy()

    # ------------------------------------------------------------------ kernels
    def estep(self, alpha=None, beta=None):
        fill = float('-inf') if self.log else 0.0
        st = self._stream()
        _lib.check(self.lib.mwd_fill_f64(_ptr(self.part_init), self.part_init.numel(), fill, st))
        _lib.check(self.lib.mwd_fill_f64(_ptr(self.part_trans), self.part_trans.numel(), fill, st))
        prob = self._problem(alpha, beta)
        _lib.check(self.lib.mwd_hmm_estep(C.byref(prob), st))
        _lib.check(self.lib.mwd_hmm_reduce(C.byref(prob), _ptr(self.post_idx), _ptr(self.post_off),
                                           _ptr(self.counts), st))

    def allreduce(self):
        fixed_order_allreduce(self.counts, self.pg, log_domain=self.log, ll_index=self.counts_len - 1)

    def mstep(self):
        a = HmmMstepArgs()
        a.log_domain = 1 if self.log else 0
        a.n_tgt_types, a.n_src_types = self.Vt, self.Vf
        a.n_lens, a.lens = len(self._lens), _np_ptr(self._lens)
        a.counts, a.acc = _ptr(self.counts), _ptr(self.acc)
        a.init, a.trans, a.obs = _ptr(self.init_t), _ptr(self.trans_t), _ptr(self.obs)
        _lib.check(self.lib.mwd_hmm_mstep(C.byref(a), self._stream()))

    def em_iteration(self):
        """E-step + reduction + (all-reduce) + M-step.  Returns the device scalar sum over all ranks
        of log p(f|e) under the parameters that ENTERED the iteration."""
        self.estep()
        self.allreduce()
        ll = self.counts[self.counts_len - 1].clone()
        self.mstep()
        return ll

    def loglik_sum(self):
        self.estep()
        self.allreduce()
        return self.counts[self.counts_len - 1].clone()

    def align(self, unk_prob=10e-12):
        torch = self.torch
        pk = self.pk
        ali = torch.empty((max(len(pk.src), 1),), dtype=torch.int32, device=self.device)
        ap = torch.zeros((max(int(pk.ap_off[-1]), 1),), dtype=torch.float64, device=self.device)
        prob = self._problem()
        _lib.check(self.lib.mwd_hmm_align(C.byref(prob), float(unk_prob), _ptr(ali), _ptr(ap),
                                          _ptr(self.ap_off), self._stream()))
        return ali[:len(pk.src)], ap[:int(pk.ap_off[-1])]

    def dense_sweeps(self):
        """forward()/backward() values for every slot of the shard (alpha, beta)."""
        torch = self.torch
        al = torch.zeros((max(self.pk.n_slots, 1),), dtype=torch.float64, device=self.device)
        be = torch.zeros_like(al)
        self.estep(al, be)
        return al, be


class SegmentHMMEngine(PlainHMMEngine):
    """Log-domain HMM over [NULL]+concepts with diagonal-Gaussian(-mixture) emissions on segment
    embeddings (the acoustic model hmm/audio_segembed_hmm_word_discoverer.py was written against).
    Source "tokens" are rows of the embedding matrix; everything else reuses the plain-state kernels
    with dense emissions."""

    def __init__(self, tgt_ids, embeddings, n_tgt_types, n_mix, device=None, rank=0, world=1,
                 emb_dtype=np.float32):
        seg_ids = [np.zeros(len(x), dtype=np.int32) for x in embeddings]
        pk = PackedSentences(tgt_ids, seg_ids, 1, rank=rank, world=world)
        PlainHMMEngine.__init__(self, pk, n_tgt_types, 1, True, device=device)
        torch = self.torch
        dev, f64 = self.device, torch.float64
        self.Vf = 0                                   # no discrete observation table on the device path
        self.counts_len = int(self.lib.mwd_hmm_counts_len(self.Vt, 0))
        self.counts = torch.zeros((self.counts_len,), dtype=f64, device=dev)
        self.reset_accumulators()
        self.M = int(n_mix)
        self.D = int(embeddings[0].shape[1])
        rows = np.concatenate([np.asarray(embeddings[i]).reshape(-1, self.D) for i in pk.order], axis=0)
        self.emb = torch.from_numpy(np.ascontiguousarray(rows, dtype=emb_dtype)).to(dev)
        self.emb_is_f64 = 1 if emb_dtype == np.float64 else 0
        self.row_pair = torch.from_numpy(pk.row_pair).to(dev)
        self.slot_row = torch.from_numpy(pk.slot_row).to(dev)
        self.emis = torch.zeros((max(pk.n_slots, 1),), dtype=f64, device=dev)
        self.resp = torch.zeros((max(pk.n_slots, 1), self.M), dtype=f64, device=dev)
        self.lprior = torch.zeros((self.Vt, self.M), dtype=f64, device=dev)
        self.means = torch.zeros((self.Vt, self.M, self.D), dtype=f64, device=dev)
        self.var = torch.ones((self.Vt, self.M, self.D), dtype=f64, device=dev)
        self.lnorm = torch.zeros((self.Vt, self.M), dtype=f64, device=dev)
        self.stats = torch.zeros((self.Vt, self.M, 1 + 2 * self.D), dtype=f64, device=dev)

    def set_emission_params(self, lprior, means, var):
        torch = self.torch
        self.lprior.copy_(torch.from_numpy(np.ascontiguousarray(lprior, dtype=np.float64)))
        self.means.copy_(torch.from_numpy(np.ascontiguousarray(means, dtype=np.float64)))
        self.var.copy_(torch.from_numpy(np.ascontiguousarray(var, dtype=np.float64)))

    def set_chain_params(self, init, trans):
        it, tt = tables_to_dense(init, trans)
        self.init_t.copy_(self.torch.from_numpy(it))
        self.trans_t.copy_(self.torch.from_numpy(tt))

    def get_all_params(self):
        init, trans = dense_to_tables(self.init_t.cpu().numpy(), self.trans_t.cpu().numpy(), self.pk.lens)
        return init, trans, self.lprior.cpu().numpy(), self.means.cpu().numpy(), self.var.cpu().numpy()

    def emission(self):
        prob = self._problem()
        _lib.check(self.lib.mwd_hmm_gauss_emission(C.byref(prob), _ptr(self.emb), self.emb_is_f64, self.D, self.M,
                                                   _ptr(self.lprior), _ptr(self.means), _ptr(self.var),
                                                   _ptr(self.lnorm), _ptr(self.emis), _ptr(self.resp), self._stream()))

    def estep(self, alpha=None, beta=None):
        self.emission()
        PlainHMMEngine.estep(self, alpha, beta)

    def em_iteration(self, update_var=False):
        self.estep()
        self.allreduce()
        ll = self.counts[self.counts_len - 1].clone()
        self.mstep()
        prob = self._problem()
        st = self._stream()
        _lib.check(self.lib.mwd_hmm_gauss_stats(C.byref(prob), _ptr(self.emb), self.emb_is_f64, self.D, self.M,
                                                _ptr(self.resp), _ptr(self.post_idx), _ptr(self.post_off),
                                                _ptr(self.stats), st))
        fixed_order_allreduce(self.stats, self.pg)
        _lib.check(self.lib.mwd_hmm_gauss_update(self.Vt, self.M, self.D, _ptr(self.stats), 1 if update_var else 0,
                                                 _ptr(self.lprior), _ptr(self.means), _ptr(self.var), st))
        return ll

    def align(self, unk_prob=10e-12):
        self.emission()
        return PlainHMMEngine.align(self, unk_prob)
