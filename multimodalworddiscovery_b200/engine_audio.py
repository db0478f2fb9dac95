"""Device engine of the dense-emission class ``ImageAudioHMMWordDiscoverer`` (SURVEY 8 f2,
reference hmm_dnn/image_audio_hmm_word_discoverer.py).

The model is the image-phone (region i, concept k)-state HMM whose discrete emission ``obs[:, x_t]``
is replaced by ``E[t] = phoneProbs @ softmaxLayerA(a_t)`` (:286-288).  The recursion, concept-chain
and Viterbi kernels therefore run unchanged on ``obsT = E`` (frames x K) with the identity phone
sequence; this engine adds the frame posterior / dense emission GEMMs in front and the
concept-phone count reduction (updateConceptPhoneCounts :486-493, phoneCounts :230-231) behind.
No CPU fallback, like ``IKEngine``.
"""
import ctypes as C

import numpy as np

from . import _lib
from .corpus import pack_pairs
from .engine import IKEngine, _ptr


def pack_audio_pairs(feats_list, audio_list, feat_dtype=np.float32, rank=0, world=1):
    """Sort by (n regions, T frames) and shard like ``pack_pairs``; returns (packed, audio) where
    ``audio`` is the (frames, Da) float64 matrix in packed order and ``packed.phones`` is the identity
    frame index (the kernels read emission row ``phones[t]``)."""
    pk = pack_pairs(feats_list, [np.zeros(len(a), dtype=np.int32) for a in audio_list], feat_dtype=feat_dtype,
                    rank=rank, world=world)
    Da = audio_list[0].shape[1] if len(audio_list) else 0
    rows = [np.asarray(audio_list[int(ex)], dtype=np.float64).reshape(-1, Da) for ex in pk.order]
    audio = np.concatenate(rows, axis=0) if rows else np.zeros((0, Da))
    pk.phones = np.arange(pk.n_phones_total, dtype=np.int32)
    return pk, np.ascontiguousarray(audio)


class IKAudioEngine(IKEngine):
    """``gaussian=True`` is ImageAudioGaussianHMMWordDiscoverer: RBF posteriors on both sides (``post`` =
    musV (K, D), ``WA`` = musA (nPhones, Da)) and NO EPS floor in the E- and M-step."""

    def __init__(self, packed, audio, n_concepts, n_phones, device=None, process_group=None, gaussian=False):
        self._audio_ready = False
        self.nPh = int(n_phones)
        if packed.n_phones_total * int(n_concepts) >= 2 ** 31:
            raise ValueError('dense emission table of %d frames x %d concepts exceeds 2^31 entries per shard'
                             % (packed.n_phones_total, n_concepts))
        IKEngine.__init__(self, packed, n_concepts, n_phones, gaussian=bool(gaussian), device=device,
                          keep_concept_counts_a=True, process_group=process_group)
        if self.gaussian:
            self._mstep_extra_flags = 8           # MWD_MSTEP_NO_FLOORS
        torch, dev, f64 = self.torch, self.device, self.torch.float64
        Tt = max(packed.n_phones_total, 1)
        self.Da = int(audio.shape[1])
        self.afeats = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float64)).to(dev)
        acols = self.Da if self.gaussian else self.Da + 1
        self.WA = torch.zeros((self.nPh, acols), dtype=f64, device=dev)
        self.gradA0 = torch.zeros((self.nPh, acols), dtype=f64, device=dev)
        self.w_scratchA = torch.zeros((self.nPh, self.Da + 1), dtype=f64, device=dev) if self.gaussian else None
        self.PH = torch.empty((Tt, self.nPh), dtype=f64, device=dev)
        self.E = torch.empty((Tt, self.K), dtype=f64, device=dev)
        self.v_scratch = torch.empty((self.K, self.nPh + 1), dtype=f64, device=dev)
        self.cpc_partials = torch.empty((int(self.lib.mwd_concept_phone_partials_len(self.K, self.nPh)),),
                                        dtype=f64, device=dev)
        self._audio_ready = True

    # the kernels see the dense emission table as a (frames x K) obsT and no phone-count table
    def _problem(self, with_cA=True):
        p = IKEngine._problem(self, with_cA=with_cA)
        p.n_phone_types = max(self.pk.n_phones_total, 1)
        p.part_phone = C.c_void_p(0)
        p.no_floor = 1 if self.gaussian else 0
        if self._audio_ready:
            p.obsT = _ptr(self.E)
        return p

    def _param_tensors(self):
        # the audio posterior weights are model parameters too (snapshot / restore, CUDA-graph warm-up)
        ts = IKEngine._param_tensors(self)
        if self._audio_ready:
            ts.append(self.WA)
        return ts

    def set_audio_param(self, WA):
        WA = np.ascontiguousarray(np.asarray(WA, dtype=np.float64))
        if WA.shape != tuple(self.WA.shape):
            raise ValueError('WA shape %s != %s' % (WA.shape, tuple(self.WA.shape)))
        self.WA.copy_(self.torch.from_numpy(WA))

    def get_audio_param(self):
        return self.WA.cpu().numpy().copy()

    def posterior(self, width=1.0):
        """pz (softmaxLayerV :543-547), frame posteriors (softmaxLayerA :549-554) and E (:288)."""
        IKEngine.posterior(self, width)
        lib, st = self.lib, self._stream()
        Tt = self.pk.n_phones_total
        self._frame_posterior(self.afeats, Tt, self.PH, width)
        _lib.check(lib.mwd_dense_emission(_ptr(self.PH), _ptr(self.obsT), Tt, self.nPh, self.K,
                                          _ptr(self.v_scratch), _ptr(self.E), st))

    def _frame_posterior(self, a_d, T, out, width):
        lib, st = self.lib, self._stream()
        if self.gaussian:    # softmaxLayerA of the gaussian class (:573-583)
            _lib.check(lib.mwd_posterior_gaussian(_ptr(a_d), 1, T, self.Da, _ptr(self.WA), float(width), self.nPh,
                                                  _ptr(self.w_scratchA), _ptr(out), st))
        else:
            _lib.check(lib.mwd_posterior_linear(_ptr(a_d), 1, T, self.Da, _ptr(self.WA), self.nPh, _ptr(out), st))

    def estep(self, width=1.0, with_cA=True, timers=None):
        lib, st = self.lib, self._stream()
        self._zero_partials()
        self.posterior(width)
        prob = self._problem(with_cA=True)
        _lib.check(lib.mwd_ik_estep(C.byref(prob), st))
        self._ca_valid = True
        self._cA_fresh = True
        _lib.check(lib.mwd_ik_concept_counts(C.byref(prob), st))
        red = self._problem(with_cA=True)
        red.n_phone_types = self.nPh                  # layout of `counts`: phone block is nPhones x K
        _lib.check(lib.mwd_ik_reduce_counts(C.byref(red), _ptr(self.counts), st))
        _lib.check(lib.mwd_concept_phone_counts(_ptr(self.cA), _ptr(self.PH), self.pk.n_phones_total, self.K,
                                                self.nPh, _ptr(self.cpc_partials), _ptr(self.counts), st))
        _lib.check(lib.mwd_ik_posterior_grad(C.byref(prob), _ptr(self.grad_partials), _ptr(self.grad), st))
        if self.gaussian:
            # argmax_ph of this E-step's frame posteriors (= the 'phone_clusters' of printAlignment, which the
            # reference derives from the stored conceptPhoneCounts): kept because PH is overwritten by the
            # next posterior() call (log-likelihood / decode under the updated parameters)
            Tt = self.pk.n_phones_total
            if getattr(self, 'phone_clusters', None) is None:
                self.phone_clusters = self.torch.empty((max(Tt, 1),), dtype=self.torch.int32, device=self.device)
            _lib.check(lib.mwd_argmax_rows(_ptr(self.PH), Tt, self.nPh, _ptr(self.phone_clusters), st))

    def mstep(self, lr, momentum, width=1.0, freeze_trans=False):
        IKEngine.mstep(self, lr, momentum, width, freeze_trans)
        # updateSoftmaxWeightA (:527-541): Delta = sum_k conceptPhoneCount - phProb vanishes identically
        # (the per-frame normalised outer product summed over k IS phProb), the reference's dW is
        # rounding noise <= 1e-17 (tests/golden/ia_*.npz) -> WA <- (1 - momentum) * WA
        _lib.check(self.lib.mwd_sgd_update(_ptr(self.WA), _ptr(self.gradA0), self.WA.numel(), 1.0, float(lr),
                                           float(momentum), self._stream()))

    # ------------------------------------------------------------------ single-pair API
    def emission_rows(self, a, width=1.0):
        """(softmaxLayerA(aSen), probs_x_given_z) of an arbitrary (T, Da) frame block: device tensors."""
        torch = self.torch
        a_d = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
        T = int(a_d.shape[0])
        ph = torch.empty((max(T, 1), self.nPh), dtype=torch.float64, device=self.device)
        E = torch.empty((max(T, 1), self.K), dtype=torch.float64, device=self.device)
        st = self._stream()
        self._frame_posterior(a_d, T, ph, width)
        _lib.check(self.lib.mwd_dense_emission(_ptr(ph), _ptr(self.obsT), T, self.nPh, self.K,
                                               _ptr(self.v_scratch), _ptr(E), st))
        return ph[:T], E[:T]

    def decode_pair_audio(self, v, a, alignment=None, width=1.0):
        _, E = self.emission_rows(a, width)
        return self.decode_pair(v, np.arange(len(a), dtype=np.int32), floor_norm=True, width=width,
                                alignment=alignment, obsT=E)

    def dense_sweep_audio(self, v, a, backward=False, width=1.0):
        _, E = self.emission_rows(a, width)
        pz = self.posterior_rows(np.asarray(v), width)
        return self.dense_sweep(pz, np.arange(len(a), dtype=np.int32), backward=backward, obsT=E)
