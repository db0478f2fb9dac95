"""Device state and the one-EM-iteration driver of the (region i, concept k)-state model.

PyTorch is plumbing only (device memory, streams, torch.distributed); every computation is a
kernel of libmwd_b200.so called through the C ABI of include/mwd_b200.h.  There is no CPU
fallback: without a CUDA device or without the built library, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import IkMstepArgs, IkProblem, MwdError, NMAX, PartialSizes
from .corpus import dense_to_tables, tables_to_dense
from .dist import fixed_order_allreduce, is_distributed


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


class IKEngine(object):
    """Holds one rank's packed shard, the model parameters and all workspaces in HBM."""

    def __init__(self, packed, n_concepts, n_phone_types, gaussian=False, device=None,
                 keep_concept_counts_a=False, process_group=None, hidden_dim=0, mixed_precision=0):
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
        self.K, self.P = int(n_concepts), int(n_phone_types)
        if not (1 <= self.K <= _lib.KMAX):
            raise ValueError('n_words=%d outside [1,%d]' % (self.K, _lib.KMAX))
        self.gaussian = bool(gaussian)
        # MWD_MIXED_* bits (include/mwd_b200.h): 0 = float64 everywhere (reference arithmetic, default)
        self.mixed = _lib.mixed_bits(mixed_precision)
        self.H = int(hidden_dim)              # > 0: two-layer (ReLU MLP) image posterior
        self.two_layer = self.H > 0
        self.D = int(packed.feats.shape[1])
        self.pg = process_group
        self.geom = _lib.geometry()
        dev = self.device
        f64 = torch.float64

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=False)

        # corpus
        self.region_off = up(packed.region_off)
        self.phone_off = up(packed.phone_off)
        self.feats = up(packed.feats)
        self.phones = up(packed.phones)
        self.feat_is_f64 = 1 if packed.feats.dtype == np.float64 else 0
        # host-side bucket descriptors (kept alive for the struct)
        self._bucket_n = np.ascontiguousarray(packed.bucket_n, dtype=np.int32)
        self._bucket_lo = np.ascontiguousarray(packed.bucket_lo, dtype=np.int64)
        self._bucket_tmax = np.ascontiguousarray(packed.bucket_tmax, dtype=np.int32)
        N, R, Tt = packed.n_pairs, packed.n_regions, packed.n_phones_total
        # parameters
        self.init_t = torch.zeros((NMAX + 1, NMAX), dtype=f64, device=dev)
        self.trans_t = torch.zeros((NMAX + 1, NMAX * NMAX), dtype=f64, device=dev)
        self.obsT = torch.zeros((self.P, self.K), dtype=f64, device=dev)
        pcols = self.D if self.gaussian else (self.H + 1 if self.two_layer else self.D + 1)
        self.post = torch.zeros((self.K, pcols), dtype=f64, device=dev)
        if self.two_layer:
            self.V_t = torch.zeros((self.H, self.D + 1), dtype=f64, device=dev)
            self.hidden = torch.empty((max(R, 1), self.H), dtype=f64, device=dev)
            self.eps = torch.empty((max(R, 1), self.H), dtype=f64, device=dev)
        self.w_scratch = torch.zeros((self.K, self.D + 1), dtype=f64, device=dev) if self.gaussian else None
        # per-region / per-pair outputs
        self.pz = torch.empty((max(R, 1), self.K), dtype=f64, device=dev)
        self.cC = torch.empty((max(R, 1), self.K), dtype=f64, device=dev)
        self.pair_ll = torch.zeros((max(N, 1),), dtype=f64, device=dev)
        # conceptCountsA (Ttot x K float64, 26 KB per MSCOCO pair) is only materialised on request
        # (materialize_cA); what printAlignment needs of it -- its row argmax -- is written by the
        # E-step kernel itself into `ca` (4 bytes per phone)
        self.cA = torch.empty((max(Tt, 1), self.K), dtype=f64, device=dev) if keep_concept_counts_a else None
        self.ca = torch.empty((max(Tt, 1),), dtype=torch.int32, device=dev)
        self._ca_valid = False
        self._prev = None      # parameters that ENTERED the last EM iteration (device snapshot)
        # partial tables + reduced buffer [counts | grad] (the all-reduce unit)
        ps = PartialSizes()
        _lib.check(self.lib.mwd_ik_partial_sizes(self.K, self.P, C.byref(ps)))
        self.part = torch.zeros((ps.phone_elems + ps.init_elems + ps.trans_elems,), dtype=f64, device=dev)
        self._part_phone = self.part[:ps.phone_elems]
        self._part_init = self.part[ps.phone_elems:ps.phone_elems + ps.init_elems]
        self._part_trans = self.part[ps.phone_elems + ps.init_elems:]
        self.counts_len = int(self.lib.mwd_ik_counts_len(self.K, self.P))
        if self.two_layer:
            self.grad_len = self.K * (self.H + 1)
            self.gradV_len = self.H * (self.D + 1)
        else:
            self.grad_len = self.K * (self.D + 1)
            self.gradV_len = 0
        self.reduced = torch.zeros((self.counts_len + self.grad_len + self.gradV_len,), dtype=f64, device=dev)
        self.counts = self.reduced[:self.counts_len]
        self.grad = self.reduced[self.counts_len:self.counts_len + self.grad_len]
        self.gradV = self.reduced[self.counts_len + self.grad_len:]
        gp_len = int(self.lib.mwd_outer_grad_partials_len(self.K, self.D))
        if self.two_layer:
            gp_len = max(gp_len, int(self.lib.mwd_outer_grad_partials_len(self.K, self.H)),
                         int(self.lib.mwd_outer_grad_partials_len(self.H, self.D)))
        # the tensor-core (tcgen05) gradient GEMM keeps one partial table per SM: size for it up front when it can run
        if not self.two_layer and self.lib.mwd_posterior_grad_tc_supported(self.feat_is_f64, self.D, self.K):
            gp_len = max(gp_len, int(self.lib.mwd_posterior_grad_tc_partials_len(self.K, self.D)))
        self.grad_partials = torch.empty((gp_len,), dtype=f64, device=dev)
        self._lens = np.ascontiguousarray(np.array(packed.lens, dtype=np.int32))
        self.toeplitz = 1 if len(packed.lens) >= 6 else 0     # :399
        self.n_pairs_global = int(packed.n_pairs_global)
        # per-step row statistics handed from the recursion kernel to the count post-pass
        slot = packed.ap_offsets()
        self.slot_off = torch.from_numpy(slot).to(dev)
        self.stats = torch.empty((max(4 * int(slot[-1]), 1),), dtype=f64, device=dev)
        # checkpoint scratch
        self.scratch = None
        prob = self._problem()
        need = int(self.lib.mwd_ik_scratch_bytes(C.byref(prob)))
        self.scratch = torch.empty((max(need, 8) // 8 + 1,), dtype=f64, device=dev)
        self.last_counts = None
        self._tc_split_mode = int(os.environ.get('MWD_TC_SPLIT_MODE', '0'))
        self.w_split = None
        self.set_mixed(mixed_precision)
        self._sum_scratch = torch.empty((256,), dtype=f64, device=dev)
        self._ll_out = torch.zeros((2,), dtype=f64, device=dev)   # [0]: loglik_sum, [1]: LL of the last EM iteration

    def set_mixed(self, spec):
        """Select which floor-free parts leave the FP64 pipe (see _lib.mixed_bits); callable between iterations.
        The tensor-core kernels need fp32 features and are not used by the two-layer class (posterior: linear class only)."""
        self.mixed = _lib.mixed_bits(spec)
        self._tc_posterior = bool(self.mixed & _lib.MIXED_POSTERIOR) and not self.gaussian and not self.two_layer \
            and bool(self.lib.mwd_posterior_tc_supported(self.feat_is_f64, self.D, self.K))
        self._tc_grad = bool(self.mixed & _lib.MIXED_GRAD) and not self.two_layer \
            and bool(self.lib.mwd_posterior_grad_tc_supported(self.feat_is_f64, self.D, self.K))
        if self._tc_posterior and self.w_split is None:
            nb = int(self.lib.mwd_posterior_tc_scratch_bytes(self.K, self.D))
            self.w_split = self.torch.empty((nb // 4,), dtype=self.torch.float32, device=self.device)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _problem(self, with_cA=True):
        pk = self.pk
        p = IkProblem()
        p.n_pairs, p.n_regions, p.n_phones_total = pk.n_pairs, pk.n_regions, pk.n_phones_total
        p.feat_dim, p.feat_is_f64 = self.D, self.feat_is_f64
        p.n_concepts, p.n_phone_types = self.K, self.P
        p.t_max = pk.t_max
        p.n_buckets = len(self._bucket_n)
        p.bucket_n, p.bucket_lo = _np_ptr(self._bucket_n), _np_ptr(self._bucket_lo)
        p.bucket_tmax = _np_ptr(self._bucket_tmax)
        p.region_off, p.phone_off = _ptr(self.region_off), _ptr(self.phone_off)
        p.feats, p.phones = _ptr(self.feats), _ptr(self.phones)
        p.init, p.trans, p.obsT = _ptr(self.init_t), _ptr(self.trans_t), _ptr(self.obsT)
        p.pz, p.concept_counts = _ptr(self.pz), _ptr(self.cC)
        p.pair_ll = _ptr(self.pair_ll)
        p.concept_counts_a = _ptr(self.cA) if (with_cA and self.cA is not None) else C.c_void_p(0)
        p.concept_alignment = _ptr(self.ca)
        p.mixed_precision = self.mixed
        p.part_phone, p.part_init = _ptr(self._part_phone), _ptr(self._part_init)
        p.part_trans = _ptr(self._part_trans)
        p.scratch = _ptr(self.scratch)
        p.scratch_bytes = self.scratch.numel() * 8 if self.scratch is not None else 0
        p.stats, p.slot_off = _ptr(self.stats), _ptr(self.slot_off)
        return p

    # ------------------------------------------------------------------ parameters
    def set_params(self, init, trans, obs, posterior_param, hidden_param=None):
        """init/trans: reference dicts keyed by n; obs: (K, P); posterior_param: W (K, D+1) or mus (K, D)
        (two-layer: W (K, H+1) and hidden_param V (H, D+1))."""
        torch = self.torch
        it, tt = tables_to_dense(init, trans)
        self.init_t.copy_(torch.from_numpy(it))
        self.trans_t.copy_(torch.from_numpy(tt))
        obs = np.asarray(obs, dtype=np.float64)
        if obs.shape != (self.K, self.P):
            raise ValueError('obs shape %s != (%d, %d)' % (obs.shape, self.K, self.P))
        self.obsT.copy_(torch.from_numpy(np.ascontiguousarray(obs.T)))
        pp = np.ascontiguousarray(np.asarray(posterior_param, dtype=np.float64))
        if pp.shape != tuple(self.post.shape):
            raise ValueError('posterior parameter shape %s != %s' % (pp.shape, tuple(self.post.shape)))
        self.post.copy_(torch.from_numpy(pp))
        if self.two_layer:
            hv = np.ascontiguousarray(np.asarray(hidden_param, dtype=np.float64))
            if hv.shape != tuple(self.V_t.shape):
                raise ValueError('hidden weight shape %s != %s' % (hv.shape, tuple(self.V_t.shape)))
            self.V_t.copy_(torch.from_numpy(hv))

    def get_hidden_param(self):
        return self.V_t.cpu().numpy().copy()

    def get_params(self):
        it = self.init_t.cpu().numpy()
        tt = self.trans_t.cpu().numpy()
        init, trans = dense_to_tables(it, tt, self.pk.lens)
        obs = np.ascontiguousarray(self.obsT.cpu().numpy().T)
        return init, trans, obs, self.post.cpu().numpy().copy()

    # ------------------------------------------------------------------ kernels
    def posterior(self, width=1.0):
        """pz = p(z | v) for every region of the shard (softmaxLayer)."""
        lib, st = self.lib, self._stream()
        R = self.pk.n_regions
        if self.two_layer:
            _lib.check(lib.mwd_hidden_relu(_ptr(self.feats), self.feat_is_f64, R, self.D, _ptr(self.V_t), self.H,
                                           _ptr(self.hidden), st))
            _lib.check(lib.mwd_posterior_linear(_ptr(self.hidden), 1, R, self.H, _ptr(self.post), self.K,
                                                _ptr(self.pz), st))
        elif self.gaussian:
            _lib.check(lib.mwd_posterior_gaussian(_ptr(self.feats), self.feat_is_f64, R, self.D,
                                                  _ptr(self.post), float(width), self.K,
                                                  _ptr(self.w_scratch), _ptr(self.pz), st))
        else:
            self._posterior_linear(_ptr(self.feats), R, _ptr(self.pz), st)

    def _grad_partial(self, prob, accumulate, st):
        """(conceptCounts - pz)^T [V,1] of one shard (chunk) into the partial tables: float64 DMMA kernel or the
        tcgen05 split-TF32 kernel (MWD_MIXED_GRAD)."""
        if self._tc_grad:
            _lib.check(self.lib.mwd_ik_posterior_grad_tc_partial(C.byref(prob), _ptr(self.grad_partials), accumulate,
                                                                 self._tc_split_mode, st))
        else:
            _lib.check(self.lib.mwd_ik_posterior_grad_partial(C.byref(prob), _ptr(self.grad_partials), accumulate, st))

    def _grad_finish(self, st):
        fn = self.lib.mwd_posterior_grad_tc_finish if self._tc_grad else self.lib.mwd_ik_posterior_grad_finish
        _lib.check(fn(self.K, self.D, _ptr(self.grad_partials), _ptr(self.grad), st))

    def _posterior_linear(self, f_ptr, R, pz_ptr, st):
        """softmaxLayer of the linear class: float64 DMMA kernel, or (MWD_MIXED_POSTERIOR, fp32 features) the
        tcgen05 split-TF32 kernel."""
        lib = self.lib
        if self._tc_posterior:
            _lib.check(lib.mwd_posterior_linear_tc(f_ptr, R, self.D, _ptr(self.post), self.K, pz_ptr,
                                                   _ptr(self.w_split), self._tc_split_mode, st))
        else:
            _lib.check(lib.mwd_posterior_linear(f_ptr, self.feat_is_f64, R, self.D, _ptr(self.post), self.K,
                                                pz_ptr, st))

    def loglik_sum(self, width=1.0):
        """Sum over this shard of log(max(p(x|y), EPS)) under the current parameters (device scalar)."""
        self.posterior(width)
        prob = self._problem()
        st = self._stream()
        _lib.check(self.lib.mwd_ik_loglik(C.byref(prob), st))
        _lib.check(self.lib.mwd_sum_f64(_ptr(self.pair_ll), self.pk.n_pairs, _ptr(self._sum_scratch),
                                        _ptr(self._ll_out), st))
        return self._ll_out[0]

    def _zero_partials(self):
        _lib.check(self.lib.mwd_fill_f64(_ptr(self.part), self.part.numel(), 0.0, self._stream()))

    def _ll_scalar(self):
        """The reduced sum of log-likelihoods of the last E-step (a view: consume before the next one)."""
        return self.counts[self.counts_len - 1]

    def estep(self, width=1.0, with_cA=True, timers=None):
        """E-step over the shard: fills pz, pair_ll, cC, (cA) and the reduced [counts | grad].
        ``timers``: optional list that receives (kernel name, start event, end event) triples
        recorded on the launching stream (bench.py's per-kernel roofline)."""
        lib, st = self.lib, self._stream()
        torch = self.torch

        def timed(name, fn):
            if timers is None:
                return fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            timers.append((name, e0, e1))

        self._zero_partials()
        timed('posterior', lambda: self.posterior(width))
        prob = self._problem(with_cA=with_cA and self.cA is not None)
        timed('ik_estep', lambda: _lib.check(lib.mwd_ik_estep(C.byref(prob), st)))
        self._ca_valid = True
        self._cA_fresh = bool(with_cA and self.cA is not None)
        timed('ik_concept', lambda: _lib.check(lib.mwd_ik_concept_counts(C.byref(prob), st)))
        timed('reduce_counts', lambda: _lib.check(lib.mwd_ik_reduce_counts(C.byref(prob), _ptr(self.counts), st)))
        if self.two_layer:
            timed('posterior_grad', self._two_layer_grads)
        else:
            timed('posterior_grad', lambda: (self._grad_partial(prob, 0, st), self._grad_finish(st)))

    def _two_layer_grads(self):
        """updateNeuralNetWeights :504-526: dW = (cC - pz)^T [h,1];  dV = ((cC - pz) W * (h>0))^T [v,1]."""
        lib, st = self.lib, self._stream()
        R = self.pk.n_regions
        _lib.check(lib.mwd_outer_grad(_ptr(self.hidden), 1, R, self.H, _ptr(self.cC), _ptr(self.pz), self.K,
                                      _ptr(self.grad_partials), _ptr(self.grad), st))
        _lib.check(lib.mwd_backprop_hidden(_ptr(self.cC), _ptr(self.pz), _ptr(self.post), _ptr(self.hidden), R,
                                           self.K, self.H, _ptr(self.eps), st))
        _lib.check(lib.mwd_outer_grad(_ptr(self.feats), self.feat_is_f64, R, self.D, _ptr(self.eps), C.c_void_p(0),
                                      self.H, _ptr(self.grad_partials), _ptr(self.gradV), st))

    def kernel_launches_per_iteration(self, n_chunks=None):
        """Kernels of libmwd_b200.so launched by one em_iteration / em_iteration_streamed."""
        # posterior (RBF: + expansion kernel) | per bucket: recursion + count post-pass + concept chains |
        # reduce_counts: 3 tables x 2 levels + 2 log-likelihood stages | gradient GEMM + its reduction |
        # M-step: init/trans, obs, posterior parameter
        # mixed path: + the weight-split kernel of the tensor-core posterior, + the table-preparation kernel of the
        # float32 concept chains (one per call)
        post = (2 if self.gaussian else 1) + (1 if self._tc_posterior else 0)
        prep = 1 if (self.mixed & _lib.MIXED_CONCEPT) else 0
        if n_chunks is None:
            nb = int(np.count_nonzero(np.diff(self._bucket_lo) > 0))
            return post + 3 * nb + prep + 8 + 2 + 3
        per_chunk = sum(post + 3 * len(ch['bucket_n']) + prep + 1 for ch in self.plan_chunks(n_chunks))
        return per_chunk + 8 + 1 + 3

    # ------------------------------------------------------------------ parameter snapshots (device to device)
    def _param_tensors(self):
        ts = [self.init_t, self.trans_t, self.obsT, self.post]
        if self.two_layer:
            ts.append(self.V_t)
        return ts

    def _copy_params(self, dst, src):
        for d, s_ in zip(dst, src):
            d.copy_(s_)

    def _snapshot_entering(self, width):
        """Keep the parameters an EM iteration starts from: conceptCountsA of that iteration can then be
        materialised later without having been stored (materialize_cA)."""
        if self._prev is None:
            self._prev = [self.torch.empty_like(t) for t in self._param_tensors()]
        self._copy_params(self._prev, self._param_tensors())
        self._prev_width = width

    def snapshot(self):
        """simulatedAnnealing's init_prev / trans_prev / obs_prev / W_prev (:169-172) as device copies."""
        if getattr(self, '_sa_snap', None) is None:
            self._sa_snap = [self.torch.empty_like(t) for t in self._param_tensors()]
        self._copy_params(self._sa_snap, self._param_tensors())

    def restore_snapshot(self):
        self._copy_params(self._param_tensors(), self._sa_snap)

    def perturb_posterior(self, noise, scale):
        """W (or mus) += scale * noise  (:173); ``noise``: host array drawn by the caller's RNG."""
        nz = self.torch.from_numpy(np.ascontiguousarray(noise, dtype=np.float64)).to(self.device)
        if tuple(nz.shape) != tuple(self.post.shape):
            raise ValueError('noise shape %s != %s' % (tuple(nz.shape), tuple(self.post.shape)))
        _lib.check(self.lib.mwd_sgd_update(_ptr(self.post), _ptr(nz), self.post.numel(), 1.0, float(scale), 0.0,
                                           self._stream()))

    def materialize_cA(self):
        """conceptCountsA (Ttot x K) of the LAST E-step, recomputed on demand from the parameters that
        entered it (same kernels, same launch shapes -> the values the E-step would have stored)."""
        if self.cA is not None and self._ca_valid and getattr(self, '_cA_fresh', False):
            return self.cA
        if self._prev is None:
            raise MwdError('no E-step has run yet')
        torch = self.torch
        if self.cA is None:
            self.cA = torch.empty((max(self.pk.n_phones_total, 1), self.K), dtype=torch.float64, device=self.device)
        cur = [t.clone() for t in self._param_tensors()]
        self._copy_params(self._param_tensors(), self._prev)
        self.posterior(self._prev_width)
        prob = self._problem(with_cA=True)
        prob.part_phone = C.c_void_p(0)          # counts of this pass are not wanted ...
        self._zero_partials()                    # ... and init / trans partials are scratch by now
        _lib.check(self.lib.mwd_ik_estep(C.byref(prob), self._stream()))
        self._copy_params(self._param_tensors(), cur)
        self._cA_fresh = True
        return self.cA

    def _keep_ll(self):
        """Park the iteration's summed log-likelihood (device-to-device copy, no kernel)."""
        self._ll_out[1:2].copy_(self.counts[self.counts_len - 1:self.counts_len])
        return self._ll_out[1]

    def allreduce(self):
        """Sum [counts | grad] over ranks: all_gather + fixed-rank-order sum (bitwise reproducible)."""
        fixed_order_allreduce(self.reduced, self.pg)

    def mstep(self, lr, momentum, width=1.0, freeze_trans=False):
        a = IkMstepArgs()
        a.gaussian = 1 if self.gaussian else 0
        a.flags = (3 if self.two_layer else 0) | (4 if freeze_trans else 0) | getattr(self, '_mstep_extra_flags', 0)   # MWD_MSTEP_* bits
        a.n_concepts, a.n_phone_types, a.feat_dim = self.K, self.P, self.D
        a.n_lens, a.lens = len(self._lens), _np_ptr(self._lens)
        a.toeplitz = self.toeplitz
        a.n_pairs_global = self.n_pairs_global
        a.lr, a.momentum, a.width = float(lr), float(momentum), float(width)
        a.counts, a.grad = _ptr(self.counts), _ptr(self.grad)
        a.init, a.trans, a.obsT = _ptr(self.init_t), _ptr(self.trans_t), _ptr(self.obsT)
        a.posterior_param = _ptr(self.post)
        _lib.check(self.lib.mwd_ik_mstep(C.byref(a), self._stream()))
        if self.two_layer:     # W and V by plain gradient steps (:527-528)
            invN = 1.0 / float(self.n_pairs_global)
            st = self._stream()
            _lib.check(self.lib.mwd_sgd_update(_ptr(self.post), _ptr(self.grad), self.grad_len, invN, float(lr),
                                               float(momentum), st))
            _lib.check(self.lib.mwd_sgd_update(_ptr(self.V_t), _ptr(self.gradV), self.gradV_len, invN, float(lr),
                                               float(momentum), st))

    def em_iteration(self, lr, momentum, width=1.0, with_cA=True, timers=None, freeze_trans=False):
        """One epoch body of trainUsingEM.  Returns the device scalar sum of log-likelihoods
        (over ALL ranks) of the parameters that entered the iteration."""
        self._snapshot_entering(width)
        self.estep(width, with_cA, timers)
        self.allreduce()
        ll = self._keep_ll()
        self.mstep(lr, momentum, width, freeze_trans)
        return ll

    # ------------------------------------------------------------------ small corpora: one CUDA-graph replay per iteration
    GRAPH_MAX_PAIRS = 50000

    def em_iteration_auto(self, lr, momentum, width=1.0, freeze_trans=False):
        """em_iteration, replayed from a CUDA graph when the shard is small enough to be launch-bound (the reference's
        own use: 2 000 pairs x 20 epochs, ~35 kernel launches of a few microseconds each per iteration) and no
        collective is on the path.  MWD_GRAPH=0 disables it."""
        if self.pk.n_pairs > self.GRAPH_MAX_PAIRS or is_distributed(self.pg) or os.environ.get('MWD_GRAPH', '1') == '0':
            return self.em_iteration(lr, momentum, width, freeze_trans=freeze_trans)
        return self.em_iteration_graph(lr, momentum, width, freeze_trans)

    def em_iteration_graph(self, lr, momentum, width=1.0, freeze_trans=False):
        """One EM iteration as ONE graph launch.  Graphs are keyed by the by-value kernel parameters (lr, momentum,
        width, flags); every buffer the kernels touch is owned by the engine, so replays see current parameters."""
        torch = self.torch
        key = (float(lr), float(momentum), float(width), bool(freeze_trans), self.mixed, self.toeplitz,
               getattr(self, '_mstep_extra_flags', 0))
        graphs = self.__dict__.setdefault('_graphs', {})
        g = graphs.get(key)
        if g is None:
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(device=self.device)
            keep = [t.clone() for t in self._param_tensors()]
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                # un-captured pass first: first-call allocations and module loads must not happen under capture
                self.em_iteration(lr, momentum, width, with_cA=False, freeze_trans=freeze_trans)
                self._copy_params(self._param_tensors(), keep)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    self.em_iteration(lr, momentum, width, with_cA=False, freeze_trans=freeze_trans)
            cur.wait_stream(side)
            if len(graphs) >= 4:                       # lr decays by value: keep only the recent few
                graphs.pop(next(iter(graphs)))
            graphs[key] = g
        g.replay()
        self._ca_valid, self._cA_fresh, self._prev_width = True, False, width
        return self._ll_out[1]

    # ------------------------------------------------------------------ streamed (host-resident) corpus
    def plan_chunks(self, n_chunks):
        """Contiguous chunks of the (n, T)-sorted shard with their own bucket descriptors."""
        pk = self.pk
        N = pk.n_pairs
        n_chunks = max(1, min(int(n_chunks), N))
        edges = np.linspace(0, N, n_chunks + 1).astype(np.int64)
        chunks = []
        for c in range(n_chunks):
            lo, hi = int(edges[c]), int(edges[c + 1])
            if hi <= lo:
                continue
            bn, blo, btm = [], [], []
            for b in range(len(self._bucket_n)):
                s, e = max(lo, int(self._bucket_lo[b])), min(hi, int(self._bucket_lo[b + 1]))
                if e > s:
                    bn.append(int(self._bucket_n[b]))
                    blo.append(s - lo)
                    btm.append(int(self._bucket_tmax[b]))
            blo.append(hi - lo)
            chunks.append(dict(lo=lo, hi=hi, r_lo=int(pk.region_off[lo]), r_hi=int(pk.region_off[hi]),
                               p_lo=int(pk.phone_off[lo]), p_hi=int(pk.phone_off[hi]),
                               bucket_n=np.array(bn, dtype=np.int32), bucket_lo=np.array(blo, dtype=np.int64),
                               bucket_tmax=np.array(btm, dtype=np.int32)))
        return chunks

    def _chunk_problem(self, ch, for_grad=False):
        p = self._problem(with_cA=False)
        lo, hi, r_lo, r_hi = ch['lo'], ch['hi'], ch['r_lo'], ch['r_hi']
        p.n_pairs = hi - lo
        p.n_buckets = len(ch['bucket_n'])
        p.bucket_n, p.bucket_lo = _np_ptr(ch['bucket_n']), _np_ptr(ch['bucket_lo'])
        p.bucket_tmax = _np_ptr(ch['bucket_tmax'])
        # offset arrays are shifted to the chunk's first pair; their VALUES stay absolute row / phone
        # indices, so the per-region and per-phone arrays keep their base pointers
        p.region_off = _ptr(self.region_off[lo:])
        p.phone_off = _ptr(self.phone_off[lo:])
        p.pair_ll = _ptr(self.pair_ll[lo:])
        p.slot_off = _ptr(self.slot_off[lo:])
        if for_grad:   # the gradient GEMM walks rows [0, n_regions) of the arrays it is given
            p.n_regions = r_hi - r_lo
            p.feats = _ptr(self.feats[r_lo:r_hi])
            p.pz = _ptr(self.pz[r_lo:r_hi])
            p.concept_counts = _ptr(self.cC[r_lo:r_hi])
        return p

    def em_iteration_streamed(self, host, lr, momentum, width=1.0, n_chunks=16):
        """One EM iteration with the corpus streamed from (pinned) host memory: the copy of chunk
        c+1 overlaps the kernels of chunk c.  ``host``: dict of pinned CPU tensors region_off,
        phone_off, feats, phones laid out like the device buffers.  Same result as em_iteration
        up to the association order of the chunked gradient partials."""
        torch, lib = self.torch, self.lib
        if self.two_layer:
            raise MwdError('em_iteration_streamed does not support the two-layer posterior yet')
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        main = torch.cuda.current_stream(self.device)
        if getattr(self, '_chunks', None) is None or len(self._chunks) != n_chunks:
            self._chunks = self.plan_chunks(n_chunks)
        chunks = self._chunks
        cs.wait_stream(main)                       # previous iteration no longer reads the buffers
        events = []
        with torch.cuda.stream(cs):
            self.region_off.copy_(host['region_off'], non_blocking=True)
            self.phone_off.copy_(host['phone_off'], non_blocking=True)
            for ch in chunks:
                self.feats[ch['r_lo']:ch['r_hi']].copy_(host['feats'][ch['r_lo']:ch['r_hi']], non_blocking=True)
                self.phones[ch['p_lo']:ch['p_hi']].copy_(host['phones'][ch['p_lo']:ch['p_hi']], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                events.append(ev)
        st = self._stream()
        self._zero_partials()
        for c, (ch, ev) in enumerate(zip(chunks, events)):
            main.wait_event(ev)
            R = ch['r_hi'] - ch['r_lo']
            f_ptr, pz_ptr = _ptr(self.feats[ch['r_lo']:ch['r_hi']]), _ptr(self.pz[ch['r_lo']:ch['r_hi']])
            if self.gaussian:
                _lib.check(lib.mwd_posterior_gaussian(f_ptr, self.feat_is_f64, R, self.D, _ptr(self.post),
                                                      float(width), self.K, _ptr(self.w_scratch), pz_ptr, st))
            else:
                self._posterior_linear(f_ptr, R, pz_ptr, st)
            prob = self._chunk_problem(ch)
            _lib.check(lib.mwd_ik_estep(C.byref(prob), st))
            _lib.check(lib.mwd_ik_concept_counts(C.byref(prob), st))
            gprob = self._chunk_problem(ch, for_grad=True)
            self._grad_partial(gprob, 1 if c > 0 else 0, st)
        full = self._problem(with_cA=False)
        _lib.check(lib.mwd_ik_reduce_counts(C.byref(full), _ptr(self.counts), st))
        self._grad_finish(st)
        self.allreduce()
        ll = self._keep_ll()
        self.mstep(lr, momentum, width)
        return ll

    def _floor_flags(self, floor_norm):
        # bit 0: floored alignProbs normaliser; bit 1: un-floored Viterbi scores (two-layer class)
        return (1 if (floor_norm or self.two_layer) else 0) | (2 if self.two_layer else 0)

    def decode(self, floor_norm=False, want_probs=True, width=1.0, want_cluster_scores=False):
        """align + cluster for every pair of the shard under the CURRENT parameters.
        Returns (alignment int32 (Ttot,), image_concepts int32 (R,), align_probs f64 ragged | None)."""
        torch = self.torch
        pk = self.pk
        self.posterior(width)
        ali = torch.empty((max(pk.n_phones_total, 1),), dtype=torch.int32, device=self.device)
        ic = torch.empty((max(pk.n_regions, 1),), dtype=torch.int32, device=self.device)
        ap = ap_off = None
        if want_probs:
            off = pk.ap_offsets()
            ap_off = torch.from_numpy(off).to(self.device)
            ap = torch.empty((max(int(off[-1]), 1),), dtype=torch.float64, device=self.device)
        prob = self._problem()
        cs = None
        if want_cluster_scores:
            cs = torch.empty((max(pk.n_regions, 1), self.K), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.mwd_ik_decode(C.byref(prob), self._floor_flags(floor_norm), 0, _ptr(ali), _ptr(ap),
                                          _ptr(ap_off), _ptr(ic), _ptr(cs), self._stream()))
        if want_cluster_scores:
            return ali[:pk.n_phones_total], ic[:pk.n_regions], ap, cs[:pk.n_regions]
        return ali[:pk.n_phones_total], ic[:pk.n_regions], ap

    def decode_pair(self, v, x, floor_norm=False, width=1.0, alignment=None, obsT=None):
        """align()/cluster() of ONE arbitrary pair under the current parameters.
        Returns (alignment (T,), align_probs (T, n), image_concepts (n,), cluster_scores (n, K)).
        ``obsT``: optional device emission table (rows x K) used instead of the model's (the
        dense-emission classes pass the pair's E and x = arange(T))."""
        torch = self.torch
        dev = self.device
        n, T = int(v.shape[0]), int(len(x))
        if not (1 <= n <= NMAX):
            raise ValueError('n=%d outside [1,%d]' % (n, NMAX))
        # single-pair API: features go up as float64 whatever the corpus storage type is (no rounding)
        v_d = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev)
        x_d = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).to(dev)
        roff = torch.tensor([0, n], dtype=torch.int32, device=dev)
        poff = torch.tensor([0, T], dtype=torch.int32, device=dev)
        pz = torch.empty((n, self.K), dtype=torch.float64, device=dev)
        st = self._stream()
        self._posterior_of(v_d, n, pz, width, is64=1)
        p = IkProblem()
        p.n_pairs, p.n_regions, p.n_phones_total = 1, n, T
        p.feat_dim, p.feat_is_f64, p.n_concepts, p.n_phone_types = self.D, 1, self.K, self.P
        p.t_max, p.n_buckets = T, 0
        p.region_off, p.phone_off, p.feats, p.phones = _ptr(roff), _ptr(poff), _ptr(v_d), _ptr(x_d)
        p.init, p.trans, p.obsT, p.pz = _ptr(self.init_t), _ptr(self.trans_t), _ptr(self.obsT), _ptr(pz)
        if obsT is not None:
            p.obsT, p.n_phone_types = _ptr(obsT), int(obsT.shape[0])
        given = alignment is not None
        if given:
            ali = torch.from_numpy(np.ascontiguousarray(alignment, dtype=np.int32)).to(dev)
        else:
            ali = torch.empty((T,), dtype=torch.int32, device=dev)
        ap = torch.zeros((T * n,), dtype=torch.float64, device=dev)
        ap_off = torch.tensor([0, T * n], dtype=torch.int64, device=dev)
        ic = torch.empty((n,), dtype=torch.int32, device=dev)
        cs = torch.empty((n, self.K), dtype=torch.float64, device=dev)
        _lib.check(self.lib.mwd_ik_decode(C.byref(p), self._floor_flags(floor_norm), 1 if given else 0, _ptr(ali),
                                          _ptr(ap), _ptr(ap_off), _ptr(ic), _ptr(cs), st))
        return (ali.cpu().numpy(), ap.cpu().numpy().reshape(T, n), ic.cpu().numpy(), cs.cpu().numpy())

    def concept_alignment(self):
        """argmax_k conceptCountsA[t][k] for every phone of the shard (printAlignment :628): written by
        the last E-step itself, no conceptCountsA needed."""
        if not self._ca_valid:
            raise MwdError('no E-step has run yet')
        return self.ca[:self.pk.n_phones_total]

    def concept_alignment_from_cA(self):
        """Same quantity from a materialised conceptCountsA (cross-check of the fused argmax)."""
        cA = self.materialize_cA()
        torch = self.torch
        Tt = self.pk.n_phones_total
        out = torch.empty((max(Tt, 1),), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.mwd_argmax_rows(_ptr(cA), Tt, self.K, _ptr(out), self._stream()))
        return out[:Tt]

    def dense_sweep(self, pz_pair, phones_pair, backward=False, obsT=None):
        """forward()/backward() of one pair: (T, n, K) tensor under the current parameters."""
        torch = self.torch
        obsT = self.obsT if obsT is None else obsT
        pz_d = torch.from_numpy(np.ascontiguousarray(pz_pair, dtype=np.float64)).to(self.device)
        ph_d = torch.from_numpy(np.ascontiguousarray(phones_pair, dtype=np.int32)).to(self.device)
        n, K = pz_pair.shape
        T = len(phones_pair)
        out = torch.zeros((T, n, K), dtype=torch.float64, device=self.device)
        if backward:
            _lib.check(self.lib.mwd_ik_backward_dense(_ptr(pz_d), _ptr(ph_d), T, n, K, _ptr(self.trans_t),
                                                      _ptr(obsT), _ptr(out), self._stream()))
        else:
            _lib.check(self.lib.mwd_ik_forward_dense(_ptr(pz_d), _ptr(ph_d), T, n, K, _ptr(self.init_t),
                                                     _ptr(self.trans_t), _ptr(obsT), _ptr(out),
                                                     self._stream()))
        return out.cpu().numpy()

    def posterior_rows(self, v, width=1.0):
        """softmaxLayer(vSen) for an arbitrary (n, D) feature block under the current parameters."""
        torch = self.torch
        v_d = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(self.device)
        out = torch.empty((v.shape[0], self.K), dtype=torch.float64, device=self.device)
        self._posterior_of(v_d, v.shape[0], out, width, is64=1)
        return out.cpu().numpy()

    def _posterior_of(self, v_d, n, out, width=1.0, is64=None):
        """softmaxLayer of an arbitrary device feature block under the current parameters."""
        lib, st = self.lib, self._stream()
        is64 = self.feat_is_f64 if is64 is None else is64
        if self.two_layer:
            h = self.torch.empty((n, self.H), dtype=self.torch.float64, device=self.device)
            _lib.check(lib.mwd_hidden_relu(_ptr(v_d), is64, n, self.D, _ptr(self.V_t), self.H, _ptr(h), st))
            _lib.check(lib.mwd_posterior_linear(_ptr(h), 1, n, self.H, _ptr(self.post), self.K, _ptr(out), st))
        elif self.gaussian:
            _lib.check(lib.mwd_posterior_gaussian(_ptr(v_d), is64, n, self.D, _ptr(self.post),
                                                  float(width), self.K, _ptr(self.w_scratch), _ptr(out), st))
        else:
            _lib.check(lib.mwd_posterior_linear(_ptr(v_d), is64, n, self.D, _ptr(self.post),
                                                self.K, _ptr(out), st))
