"""Shared host-side logic of the two (region i, concept k)-state word discoverers.

Mirrors the public surface of
  hmm_dnn/image_phone_hmm_word_discoverer.py            (ImagePhoneHMMWordDiscoverer)
  hmm_dnn/image_phone_gaussian_hmm_word_discoverer.py   (ImagePhoneGaussianHMMWordDiscoverer)
-- same constructor signatures, method names, public attributes, prints and on-disk formats --
while every per-caption computation runs as a CUDA kernel behind include/mwd_b200.h.
No NumPy/CPU fallback exists for the kernels: without the built library or a GPU, the methods
that compute raise.
"""
import ctypes as C
import json
import math
import random
import time
from copy import deepcopy

import numpy as np

from .. import _lib
from ..corpus import pack_pairs, resolve_feature_dtype

EPS = 1e-50


class OneHotCorpus(object):
    """Lazy stand-in for the reference's ``aCorpus`` (list of one-hot (T, P) float64 arrays,
    image_phone_hmm_word_discoverer.py:92-97): indexing materialises one sentence."""

    def __init__(self, phone_ids, n_types):
        self.ids = phone_ids
        self.n_types = n_types

    def __len__(self):
        return len(self.ids)

    def _one(self, x):
        a = np.zeros((len(x), self.n_types))
        a[np.arange(len(x)), x] = 1.
        return a

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(x) for x in self.ids[i]]
        return self._one(self.ids[i])

    def __iter__(self):
        for x in self.ids:
            yield self._one(x)


def one_hot_to_ids(aSen):
    """The kernels take phone ids; a reference-style one-hot (T, P) sentence collapses exactly."""
    a = np.asarray(aSen)
    if a.ndim == 1:
        return a.astype(np.int32)
    ids = np.argmax(a, axis=1)
    if not (np.all(a[np.arange(len(ids)), ids] == 1.) and np.count_nonzero(a) == len(ids)):
        raise ValueError('aSen must be one-hot rows (soft phone posteriors are the image_audio classes)')
    return ids.astype(np.int32)


def write_alignment_arrays(filePrefix, phone_off, region_off, ali, ic, ap, ap_off, concept_alignment=None,
                           concept_probs=None, cluster_probs=None, n_concepts=0, is_phoneme=True):
    """printAlignment's `.txt` + `.json` (:620-648) through the native writer of libmwd_b200.so
    (csrc/align_json.cu) from FLAT corpus-order arrays: byte-identical to the reference's per-pair
    dicts + ``json.dump(aligns, f, indent=4, sort_keys=True)``, without a Python object per number."""
    lib = _lib.load()

    def arr(a, dtype):
        return None if a is None else np.ascontiguousarray(a, dtype=dtype)

    phone_off, region_off, ap_off = arr(phone_off, np.int64), arr(region_off, np.int64), arr(ap_off, np.int64)
    ali, ic, ap = arr(ali, np.int32), arr(ic, np.int32), arr(ap, np.float64)
    ca, cp, cl = arr(concept_alignment, np.int32), arr(concept_probs, np.float64), arr(cluster_probs, np.float64)
    if ap.size != int(ap_off[-1]) or ali.size != int(phone_off[-1]) or ic.size != int(region_off[-1]):
        raise ValueError('alignment arrays do not match their offsets')

    def ptr(a):
        return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)

    _lib.check(lib.mwd_write_alignment_files((filePrefix + '.txt').encode(), (filePrefix + '.json').encode(),
                                             len(phone_off) - 1, ptr(phone_off), ptr(region_off), ptr(ali), ptr(ic),
                                             ptr(ca), ptr(ap), ptr(ap_off), ptr(cp), ptr(cl), int(n_concepts),
                                             1 if is_phoneme else 0))


def write_alignment_files(filePrefix, alis, ics, aps, concept_alignment=None, concept_probs=None,
                          cluster_probs=None, n_concepts=0, is_phoneme=True):
    """Same, from the per-pair lists the batched decode returns (corpus order)."""

    def cat(rows, dtype):
        if rows is None:
            return None
        rows = [np.asarray(r, dtype=dtype).ravel() for r in rows]
        return np.concatenate(rows) if rows else np.zeros((0,), dtype=dtype)

    T = np.array([len(a) for a in alis], dtype=np.int64)
    n = np.array([len(c) for c in ics], dtype=np.int64)
    write_alignment_arrays(filePrefix, np.concatenate([[0], np.cumsum(T)]), np.concatenate([[0], np.cumsum(n)]),
                           cat(alis, np.int32), cat(ics, np.int32), cat(aps, np.float64),
                           np.concatenate([[0], np.cumsum(T * n)]), cat(concept_alignment, np.int32),
                           cat(concept_probs, np.float64), cat(cluster_probs, np.float64), n_concepts, is_phoneme)


def ragged_to_corpus_order(flat, off, order, width=1):
    """Rows stored in packed (sorted) order -> corpus order, vectorised.  ``off``: offsets (in rows) of the
    packed pairs, ``order[s]`` = corpus index of packed pair s, ``width`` = entries per row.
    Returns (flat in corpus order, offsets in corpus order)."""
    off = np.asarray(off, dtype=np.int64)
    L = np.diff(off)
    inv = np.empty(len(order), dtype=np.int64)
    inv[np.asarray(order, dtype=np.int64)] = np.arange(len(order), dtype=np.int64)
    Lc = L[inv]
    new_off = np.concatenate([[0], np.cumsum(Lc)]).astype(np.int64)
    take = np.repeat(off[:-1][inv] - new_off[:-1], Lc) + np.arange(int(new_off[-1]), dtype=np.int64)
    flat = np.asarray(flat).reshape(-1, width) if width > 1 else np.asarray(flat)
    return flat[take], new_off


class ImagePhoneHMMBase(object):
    GAUSSIAN = False
    TWO_LAYER = False

    # ------------------------------------------------------------------ corpus
    def _read_features(self, imageFeatFile, limit=None):
        vNpz = np.load(imageFeatFile)
        keys = sorted(vNpz.keys(), key=lambda x: int(x.split('_')[-1]))       # :53
        if limit is not None:
            keys = keys[:limit]
        return [vNpz[k] for k in keys]

    def _read_captions(self, speechFeatFile, limit=None):
        """:78-97 -- phone2idx in first-seen order over the WHOLE file; ids instead of one-hots."""
        self.phone2idx = {}
        nTypes = 0
        nPhones = 0
        strs = []
        with open(speechFeatFile, 'r') as f:
            for line in f:
                aSen = line.strip().split()
                strs.append(aSen)
                for phn in aSen:
                    if phn not in self.phone2idx:
                        self.phone2idx[phn] = nTypes
                        nTypes += 1
                    nPhones += 1
        self.audioFeatDim = nTypes
        if limit is not None:
            strs = strs[:limit]
        ids = [np.array([self.phone2idx[p] for p in s], dtype=np.int32) for s in strs]
        return ids, nTypes, nPhones

    def _finish_corpus(self, ids, nTypes, nPhones):
        self._phones = ids
        self.aCorpus = OneHotCorpus(ids, nTypes)
        nImages = 0
        for ex, vfeat in enumerate(self.vCorpus):
            nImages += len(vfeat)
            if vfeat.shape[-1] == 0:                                            # :70-73
                print('ex: ', ex)
                print('vfeat empty: ', vfeat.shape)
                self.vCorpus[ex] = np.zeros((1, self.imageFeatDim))
        print('----- Corpus Summary -----')
        print('Number of examples: ', len(self.aCorpus))
        print('Number of phonetic categories: ', nTypes)
        print('Number of phones: ', nPhones)
        print('Number of objects: ', nImages)
        print("Number of word clusters: ", self.nWords)

    # ------------------------------------------------------------------ engine plumbing
    def _dist(self):
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                return dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
        return 0, 1

    def _phone_ids(self):
        a = self.aCorpus
        if isinstance(a, OneHotCorpus):
            return a.ids
        # the user replaced aCorpus with reference-style one-hot arrays
        return [one_hot_to_ids(s) for s in a]

    def _engine(self):
        """(Re)build the device engine when the corpus objects changed."""
        from ..engine import IKEngine
        token = (id(self.vCorpus), len(self.vCorpus), id(self.aCorpus), len(self.aCorpus), self.nWords,
                 self.audioFeatDim)
        if getattr(self, '_eng', None) is None or self._eng_token != token:
            rank, world = self._dist()
            dt = resolve_feature_dtype(self._feature_dtype, self.vCorpus)
            pk = pack_pairs(self.vCorpus, self._phone_ids(), feat_dtype=dt, rank=rank, world=world)
            self._eng = IKEngine(pk, self.nWords, self.audioFeatDim, gaussian=self.GAUSSIAN,
                                 device=self._device, keep_concept_counts_a=self._keep_cA,
                                 hidden_dim=(self.hiddenDim if self.TWO_LAYER else 0),
                                 mixed_precision=getattr(self, '_posterior_precision', 0))
            self._eng_token = token
            self._cA_valid = False
        return self._eng

    def _posterior_param(self):
        return self.mus if self.GAUSSIAN else self.W

    def _set_posterior_param(self, v):
        if self.GAUSSIAN:
            self.mus = v
        else:
            self.W = v

    def _push(self):
        eng = self._engine()
        eng.set_params(self.init, self.trans, self.obs, self._posterior_param(),
                       self.V if self.TWO_LAYER else None)
        return eng

    def _pull(self, eng):
        init, trans, obs, post = eng.get_params()
        for m in init:
            self.init[m] = init[m]
            self.trans[m] = trans[m]
        self.obs = obs
        self._set_posterior_param(post)
        if self.TWO_LAYER:
            self.V = eng.get_hidden_param()

    def _width(self):
        return float(getattr(self, 'width', 1.))

    # ------------------------------------------------------------------ model
    def computeTranslationLengthProbabilities(self, smoothing=None):
        """:491-521, including the per-sentence reset of lenProb[len(ts)] (only the key set and
        the last sentence's entry survive, as in the reference)."""
        for ts, fs in zip(self.vCorpus, self._phone_ids()):
            self.lenProb[len(ts)] = {}
            if len(fs) not in self.lenProb[len(ts)].keys():
                self.lenProb[len(ts)][len(fs)] = 1
            else:
                self.lenProb[len(ts)][len(fs)] += 1
        if smoothing == 'laplace':
            tLenMax = max(list(self.lenProb.keys()))
            fLenMax = max([max(list(f.keys())) for f in list(self.lenProb.values())])
            for tLen in range(tLenMax):
                for fLen in range(fLenMax):
                    if tLen not in self.lenProb:
                        self.lenProb[tLen] = {}
                        self.lenProb[tLen][fLen] = 1.
                    elif fLen not in self.lenProb[tLen]:
                        self.lenProb[tLen][fLen] = 1.
                    else:
                        self.lenProb[tLen][fLen] += 1.
        for tl in self.lenProb.keys():
            totCount = sum(self.lenProb[tl].values())
            for fl in self.lenProb[tl].keys():
                self.lenProb[tl][fl] = self.lenProb[tl][fl] / totCount

    def _load_init_trans_files(self, create_missing):
        if self.initProbFile:
            with open(self.initProbFile) as f:
                for line in f:
                    m, s, prob = line.split()
                    if create_missing and int(m) not in self.init:               # gaussian :117-118
                        self.init[int(m)] = np.zeros((int(m),))
                    self.init[int(m)][int(s)] = float(prob)
        if self.transProbFile:
            with open(self.transProbFile) as f:
                for line in f:
                    m, cur_s, next_s, prob = line.split()
                    if create_missing and int(m) not in self.trans:              # gaussian :126-127
                        self.trans[int(m)] = np.zeros((int(m), int(m)))
                    self.trans[int(m)][int(cur_s)][int(next_s)] = float(prob)

    # ------------------------------------------------------------------ EM
    def trainUsingEM(self, numIterations=20, writeModel=False, warmStart=False, convergenceEpsilon=0.01,
                     printStatus=True, debug=False, _freeze_trans=False):
        """:198-266.  The E-step (forward/backward, expected counts, concept posteriors), the
        count reduction and the M-step all run on the GPU; the per-epoch log-likelihood the
        reference computes in a separate pass (:216) is produced by the same forward sweep."""
        if not warmStart:
            self.initializeModel()
        if writeModel:
            self.printModel('initial_model.txt')
        eng = self._push()
        maxLikelihood = -np.inf
        likelihoods = np.zeros((numIterations,))
        N = len(self.vCorpus)
        for epoch in range(numIterations):
            begin_time = time.time()
            width = self._width()
            if printStatus and writeModel:
                # the reference decides about writing BEFORE the E-step, from a separate LL pass
                likelihood = float(self._allreduce_scalar(eng.loglik_sum(width))) / N
                if likelihood > maxLikelihood:
                    self._pull(eng)
                    self.printModel(self.modelName + '_iter=' + str(epoch) + '.txt')
                    # reference quirk: conceptCountsA was just reset to zeros (:213), so the
                    # intermediate dumps carry concept_alignment == 0 everywhere
                    self.printAlignment(self.modelName + '_iter=' + str(epoch) + '_alignment', debug=False,
                                        _zero_concept_alignment=True)
                    maxLikelihood = likelihood
            ll = eng.em_iteration_auto(self.lr, self.momentum, width, freeze_trans=_freeze_trans)
            self._cA_valid = True
            if printStatus:
                likelihood = float(ll) / N
                likelihoods[epoch] = likelihood
                print('Epoch', epoch, 'Average Log Likelihood:', likelihood)
            if (epoch + 1) % 10 == 0:
                self.lr /= 10
            if printStatus:
                eng.torch.cuda.synchronize(eng.device)
                print('Epoch %d takes %.2f s to finish' % (epoch, time.time() - begin_time))
        self._pull(eng)
        self._conceptCounts_cache = None
        self._conceptCountsA_cache = None
        np.save(self.modelName + '_likelihoods.npy', likelihoods)

    def _allreduce_scalar(self, t):
        rank, world = self._dist()
        if world > 1:
            import torch.distributed as dist
            t = t.clone()
            dist.all_reduce(t)
        return t

    def computeAvgLogLikelihood(self):
        """:523-531"""
        eng = self._push()
        return float(self._allreduce_scalar(eng.loglik_sum(self._width()))) / len(self.vCorpus)

    # per-pair outputs of the last E-step, reference layout (lists in corpus order)
    def _gather_rows(self, dev_rows, off):
        eng = self._eng
        rows = dev_rows.cpu().numpy()
        rank, world = self._dist()
        local = [(int(ex), rows[off[s]:off[s + 1]]) for s, ex in enumerate(eng.pk.order)]
        if world > 1:
            import torch.distributed as dist
            allp = [None] * world
            dist.all_gather_object(allp, local)
            local = [x for part in allp for x in part]
        out = [None] * len(self.vCorpus)
        for ex, r in local:
            out[ex] = r
        return out

    @property
    def conceptCounts(self):
        if getattr(self, '_conceptCounts_cache', None) is None:
            if getattr(self, '_eng', None) is None or not getattr(self, '_cA_valid', False):
                raise AttributeError("'%s' object has no attribute 'conceptCounts'" % type(self).__name__)
            self._conceptCounts_cache = self._gather_rows(self._eng.cC, self._eng.pk.region_off)
        return self._conceptCounts_cache

    @conceptCounts.setter
    def conceptCounts(self, v):
        self._conceptCounts_cache = v

    @property
    def conceptCountsA(self):
        if getattr(self, '_conceptCountsA_cache', None) is None:
            if getattr(self, '_eng', None) is None or not getattr(self, '_cA_valid', False):
                raise AttributeError("'%s' object has no attribute 'conceptCountsA'" % type(self).__name__)
            # not stored by the E-step unless modelConfigs['keep_concept_counts_a']: recomputed here from the
            # parameters that entered the last EM iteration (bit-identical, engine.materialize_cA)
            self._conceptCountsA_cache = self._gather_rows(self._eng.materialize_cA(), self._eng.pk.phone_off)
        return self._conceptCountsA_cache

    @conceptCountsA.setter
    def conceptCountsA(self, v):
        self._conceptCountsA_cache = v

    # ------------------------------------------------------------------ single-pair API
    def softmaxLayer(self, vSen, debug=False):
        """:533-541 / gaussian :501-510"""
        return self._push().posterior_rows(np.asarray(vSen), self._width())

    def forward(self, vSen, aSen, debug=False):
        """:276-304 -> (T, n, K)"""
        eng = self._push()
        pz = eng.posterior_rows(np.asarray(vSen), self._width())
        return eng.dense_sweep(pz, one_hot_to_ids(aSen), backward=False)

    def backward(self, vSen, aSen, debug=False):
        """:314-335 -> (T, n, K)"""
        eng = self._push()
        pz = eng.posterior_rows(np.asarray(vSen), self._width())
        return eng.dense_sweep(pz, one_hot_to_ids(aSen), backward=True)

    def align(self, aSen, vSen, unkProb=10e-12, debug=False):
        """:543-584 -> (bestPath list[int], alignProbs list[list[float]])"""
        eng = self._push()
        ali, ap, _, _ = eng.decode_pair(np.asarray(vSen), one_hot_to_ids(aSen), floor_norm=self.GAUSSIAN,
                                        width=self._width())
        return [int(a) for a in ali], ap.tolist()

    def cluster(self, aSen, vSen, alignment):
        """:586-597 -> (argmax list[int], scores list[list[float]])"""
        eng = self._push()
        _, _, ic, cs = eng.decode_pair(np.asarray(vSen), one_hot_to_ids(aSen), floor_norm=self.GAUSSIAN,
                                       width=self._width(), alignment=np.asarray(alignment))
        return [int(c) for c in ic], cs.tolist()

    # ------------------------------------------------------------------ I/O
    def printModel(self, fileName):
        """:599-616 (+ gaussian :629)"""
        initFile = open(fileName + '_initialprobs.txt', 'w')
        for nState in sorted(self.lenProb):
            for i in range(nState):
                initFile.write('%d\t%d\t%f\n' % (nState, i, self.init[nState][i]))
        initFile.close()
        transFile = open(fileName + '_transitionprobs.txt', 'w')
        for nState in sorted(self.lenProb):
            for i in range(nState):
                for j in range(nState):
                    transFile.write('%d\t%d\t%d\t%f\n' % (nState, i, j, self.trans[nState][i][j]))
        transFile.close()
        np.save(fileName + '_observationprobs.npy', self.obs)
        with open(fileName + '_phone2idx.json', 'w') as f:
            json.dump(self.phone2idx, f)
        if self.GAUSSIAN:
            np.save(fileName + '_visualanchors.npy', self.mus)

    def _decode_all(self, zero_concept_alignment=False):
        """Batched align + cluster + argmax(conceptCountsA) for the whole corpus, corpus order."""
        eng = self._push()
        ali, ic, ap = eng.decode(floor_norm=self.GAUSSIAN, want_probs=True, width=self._width())
        pk = eng.pk
        alis = self._gather_rows(ali, pk.phone_off)
        ics = self._gather_rows(ic, pk.region_off)
        aps = self._gather_rows(ap, pk.ap_offsets())
        if zero_concept_alignment:
            cas = [np.zeros(len(a), dtype=np.int64) for a in alis]
        else:
            # AttributeError before any trainUsingEM, as in the reference (:628)
            if not getattr(self, '_cA_valid', False):
                raise AttributeError("'%s' object has no attribute 'conceptCountsA'" % type(self).__name__)
            if getattr(self, '_conceptCountsA_cache', None) is not None:
                cas = [np.argmax(c, axis=1) for c in self._conceptCountsA_cache]
            else:
                cas = self._gather_rows(eng.concept_alignment(), pk.phone_off)
        return alis, ics, aps, cas

    def printAlignment(self, filePrefix, isPhoneme=True, debug=False, _zero_concept_alignment=False):
        """:620-648 (gaussian adds 'concept_probs', :651)"""
        rank, world = self._dist()
        if world == 1:
            return self._print_alignment_flat(filePrefix, isPhoneme, _zero_concept_alignment)
        alis, ics, aps, cas = self._decode_all(_zero_concept_alignment)
        concept_probs = None
        if self.GAUSSIAN:
            # self.conceptCounts gathers over ranks (collective): every rank must take part BEFORE the
            # non-zero ranks leave
            concept_probs = [np.asarray(c, dtype=np.float64) for c in self.conceptCounts]
        if rank != 0:
            return
        write_alignment_files(filePrefix, alis, ics, aps, concept_alignment=cas, concept_probs=concept_probs,
                              n_concepts=self.nWords, is_phoneme=isPhoneme)

    def _print_alignment_flat(self, filePrefix, isPhoneme, zero_concept_alignment):
        """Single-process fast path: batched decode -> flat arrays -> corpus order -> native writer;
        no per-pair Python objects (1 M pairs: seconds instead of the minutes of dict + json.dump)."""
        eng = self._push()
        ali, ic, ap = eng.decode(floor_norm=self.GAUSSIAN, want_probs=True, width=self._width())
        pk = eng.pk
        ali_c, phone_off = ragged_to_corpus_order(ali.cpu().numpy(), pk.phone_off, pk.order)
        ic_c, region_off = ragged_to_corpus_order(ic.cpu().numpy(), pk.region_off, pk.order)
        ap_c, ap_off = ragged_to_corpus_order(ap.cpu().numpy()[:int(pk.ap_offsets()[-1])], pk.ap_offsets(), pk.order)
        if zero_concept_alignment:
            ca_c = np.zeros(len(ali_c), dtype=np.int32)
        else:
            if not getattr(self, '_cA_valid', False):       # AttributeError before any trainUsingEM (:628)
                raise AttributeError("'%s' object has no attribute 'conceptCountsA'" % type(self).__name__)
            if getattr(self, '_conceptCountsA_cache', None) is not None:
                ca_c = np.concatenate([np.argmax(c, axis=1) for c in self._conceptCountsA_cache])
            else:
                ca_c, _ = ragged_to_corpus_order(eng.concept_alignment().cpu().numpy(), pk.phone_off, pk.order)
        cp_c = None
        if self.GAUSSIAN:
            if getattr(self, '_conceptCounts_cache', None) is not None:
                cp_c = np.concatenate([np.asarray(c, dtype=np.float64) for c in self._conceptCounts_cache]).ravel()
            else:
                cp_c, _ = ragged_to_corpus_order(eng.cC[:pk.n_regions].cpu().numpy(), pk.region_off, pk.order,
                                                 width=self.nWords)
        write_alignment_arrays(filePrefix, phone_off, region_off, ali_c, ic_c, ap_c, ap_off, concept_alignment=ca_c,
                               concept_probs=cp_c, n_concepts=self.nWords, is_phoneme=isPhoneme)

    # ------------------------------------------------------------------ simulated annealing
    def simulatedAnnealing(self, numIterations=100, T0=0.5, stepScale=5., debug=False):
        """:159-196 (gaussian :156-193).  Same accept / reject walk, same RNG draws (``np.random.normal`` for
        the jump, ``random.random`` for the Boltzmann test) and the same prints and files as the reference,
        but the model never leaves the GPU inside the loop: the reference's four ``deepcopy`` snapshots are
        device-to-device copies (``engine.snapshot / restore_snapshot``), the jump is added by a kernel, the
        inner EM iterations and the energy evaluation run on the resident tables, and the host attributes
        (``init / trans / obs / W|mus``) are refreshed only where the reference lets a caller observe them
        (``printModel`` / ``printAlignment`` on a new minimum, and on return)."""
        inner = 5 if self.GAUSSIAN else 20                                   # :174 / gaussian :171
        self.trainUsingEM(numIterations=5, warmStart=False, printStatus=True)
        eng = self._push()
        N = len(self.vCorpus)

        def energy():                                                        # -computeAvgLogLikelihood()
            return -float(self._allreduce_scalar(eng.loglik_sum(self._width()))) / N

        E0 = energy()
        Emin = E0
        count = 0
        self.sa_trace = []                                                   # (E1, E0 before the test, accepted)
        shape = self._posterior_param().shape
        for epoch in range(numIterations):
            print('Simulated Annealing Iteration %d' % epoch)
            begin_time = time.time()
            eng.snapshot()                                                   # init/trans/obs/W _prev (:169-172)
            eng.perturb_posterior(np.random.normal(size=shape), stepScale)   # :173
            # trainUsingEM(numIterations=inner, warmStart=True, printStatus=False) on the resident model (:174)
            for it in range(inner):
                eng.em_iteration_auto(self.lr, self.momentum, self._width())
                if (it + 1) % 10 == 0:
                    self.lr /= 10
            self._cA_valid = True
            self._conceptCounts_cache = None
            self._conceptCountsA_cache = None
            np.save(self.modelName + '_likelihoods.npy', np.zeros((inner,)))
            E1 = energy()
            print('Current and previous energy level: ', E1, E0)
            Tk = T0 / np.log(epoch + 2)
            if E1 > E0 and random.random() > np.exp(-(E1 - E0) / Tk):
                self.sa_trace.append((E1, E0, False))
                eng.restore_snapshot()
            else:
                self.sa_trace.append((E1, E0, True))
                if debug:
                    print('Random jump at temperature %.5f' % Tk)
                E0 = E1
                if E1 < Emin:
                    Emin = E1
                    count += 1
                    print('Update %d after %.2f s: current lowest energy level is %.5f'
                          % (count, time.time() - begin_time, Emin))
                    self._pull(eng)
                    self.printModel(self.modelName + '_%d' % count)
                    self.printAlignment(self.modelName + '_%d_alignment' % count, debug=False)
                    begin_time = time.time()
        self._pull(eng)
