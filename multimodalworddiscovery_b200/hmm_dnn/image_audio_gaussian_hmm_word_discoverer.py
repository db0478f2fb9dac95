"""Drop-in for hmm_dnn/image_audio_gaussian_hmm_word_discoverer.py
(``ImageAudioGaussianHMMWordDiscoverer``; driver: run_image2audio.py).

RBF posteriors on the image AND the audio side and -- unlike every other class of the family -- no
EPS floor anywhere in the E- or M-step (reference :241-260, :369-371, :414-415, :449-451, :629-631):
the CUDA engine runs the shared kernels with a floor value of 0 (``mwd_ik_problem.no_floor``).

Reference quirks kept on purpose: ``initializeModel`` writes ``<modelName>.txt/.json`` through
``printUnimodalCluster`` (:150); ``updateSoftmaxWeightA``'s gradient vanishes identically, so ``musA`` only
decays by ``1 - momentum``; ``is_exact=True`` raises ``NameError`` (:520 uses undefined names).
"""
import numpy as np
import math
import json
import time
from scipy.special import logsumexp
import random
from copy import deepcopy
from sklearn.cluster import KMeans

from ._ik_base import ImagePhoneHMMBase
from .image_audio_hmm_word_discoverer import ImageAudioHMMWordDiscoverer

NULL = "NULL"
DEBUG = False
EPS = 1e-50
random.seed(1)
np.random.seed(1)


class ImageAudioGaussianHMMWordDiscoverer(ImageAudioHMMWordDiscoverer):
  GAUSSIAN = True

  def __init__(self, speechFeatureFile, imageFeatureFile, modelConfigs, modelName='image_phone_hmm_word_discoverer'):
    self.modelName = modelName
    self.aCorpus = []
    self.vCorpus = []
    self.hasNull = modelConfigs.get('has_null', False)
    self.nWords = modelConfigs.get('n_words', 66)
    self.nPhones = modelConfigs.get('n_phones', 42)
    self.width = modelConfigs.get('width', 1.)
    self.momentum = modelConfigs.get('momentum', 0.)
    self.lr = modelConfigs.get('learning_rate', 10.)
    self.isExact = modelConfigs.get('is_exact', False)
    self.normalize_vfeat = modelConfigs.get('normalize_vfeat', False)
    self._device = modelConfigs.get('device', None)
    self._feature_dtype = modelConfigs.get('feature_dtype', 'auto')
    self._pair_limit = None                                   # this class reads the whole file (:66,:82)
    self.init = {}
    self.trans = {}
    self.lenProb = {}
    self.phoneProbs = None
    self.avgLogTransProb = float('-inf')
    self.readCorpus(speechFeatureFile, imageFeatureFile, debug=False)
    self.initProbFile = modelConfigs.get('init_prob_file', None)
    self.transProbFile = modelConfigs.get('trans_prob_file', None)
    self.phoneProbFile = modelConfigs.get('phone_prob_file', None)
    self.audioAnchorFile = modelConfigs.get('audio_anchor_file', None)
    self.visualAnchorFile = modelConfigs.get('visual_anchor_file', None)

  def initializeModel(self, alignments=None):
    """reference :96-150"""
    begin_time = time.time()
    self.computeTranslationLengthProbabilities()
    for m in self.lenProb:
      self.init[m] = 1. / m * np.ones((m,))
    for m in self.lenProb:
      self.trans[m] = 1. / m * np.ones((m, m))
    self._load_init_trans_files(create_missing=True)
    if self.phoneProbFile:
      self.phoneProbs = np.load(self.phoneProbFile)
    else:
      self.phoneProbs = 1. / self.nPhones * np.ones((self.nWords, self.nPhones))
    if self.visualAnchorFile:
      self.musV = np.load(self.visualAnchorFile)
      self.musA = np.load(self.audioAnchorFile)
    else:
      self.musV = KMeans(n_clusters=self.nWords).fit(np.concatenate(self.vCorpus, axis=0)).cluster_centers_
      self.musA = KMeans(n_clusters=self.nPhones).fit(np.concatenate(self.aCorpus, axis=0)).cluster_centers_
    print("Finish initialization after %0.3f s" % (time.time() - begin_time))
    self.printUnimodalCluster(filePrefix=self.modelName)

  # ------------------------------------------------------------------ engine plumbing
  def _engine(self):
    from ..engine_audio import IKAudioEngine, pack_audio_pairs
    token = (id(self.vCorpus), len(self.vCorpus), id(self.aCorpus), len(self.aCorpus), self.nWords, self.nPhones)
    if getattr(self, '_eng', None) is None or self._eng_token != token:
      rank, world = self._dist()
      from ..corpus import resolve_feature_dtype
      dt = resolve_feature_dtype(self._feature_dtype, self.vCorpus, self.aCorpus)
      pk, audio = pack_audio_pairs(self.vCorpus, self.aCorpus, feat_dtype=dt, rank=rank, world=world)
      self._eng = IKAudioEngine(pk, audio, self.nWords, self.nPhones, device=self._device, gaussian=True)
      self._eng_token = token
      self._cA_valid = False
    return self._eng

  def _push(self):
    eng = self._engine()
    eng.set_params(self.init, self.trans, self.phoneProbs, self.musV)
    eng.set_audio_param(self.musA)
    return eng

  def _pull(self, eng):
    init, trans, pp, musV = eng.get_params()
    for m in init:
      self.init[m] = init[m]
      self.trans[m] = trans[m]
    self.phoneProbs = pp
    self.musV = musV
    self.musA = eng.get_audio_param()

  def trainUsingEM(self, numIterations=20, writeModel=False, warmStart=False, convergenceEpsilon=0.01, printStatus=True, debug=False):
    if self.isExact:
      # reference :519-527 reads vSen / conceptCount / normFactor before assigning them
      raise NameError("name 'vSen' is not defined")
    ImagePhoneHMMBase.trainUsingEM(self, numIterations, writeModel, warmStart, convergenceEpsilon, printStatus, debug)

  # ------------------------------------------------------------------ single-pair API
  def softmaxLayerV(self, vSen, debug=False):
    """:561-570"""
    return self._push().posterior_rows(np.asarray(vSen), self._width())

  def softmaxLayerA(self, aSen, debug=False):
    """:572-582"""
    ph, _ = self._push().emission_rows(np.asarray(aSen), self._width())
    return ph.cpu().numpy()

  def forward(self, vSen, aSen, debug=False):
    return self._push().dense_sweep_audio(np.asarray(vSen), np.asarray(aSen), backward=False, width=self._width())

  def backward(self, vSen, aSen, debug=False):
    return self._push().dense_sweep_audio(np.asarray(vSen), np.asarray(aSen), backward=True, width=self._width())

  def align(self, aSen, vSen, unkProb=10e-12, debug=False):
    """:627-669"""
    ali, ap, _, _ = self._push().decode_pair_audio(np.asarray(vSen), np.asarray(aSen), width=self._width())
    return [int(a) for a in ali], ap.tolist()

  def cluster(self, aSen, vSen, alignment):
    """:671-683"""
    _, _, ic, cs = self._push().decode_pair_audio(np.asarray(vSen), np.asarray(aSen), alignment=np.asarray(alignment),
                                                  width=self._width())
    return [int(c) for c in ic], cs.tolist()

  @property
  def conceptPhoneCounts(self):
    """List of (T, K, nPhones) arrays of the last E-step (:275): the per-frame normalised outer products,
    rebuilt on the host from the device's normalised concept posteriors and frame posteriors."""
    if getattr(self, '_eng', None) is None or not getattr(self, '_cA_valid', False):
      raise AttributeError("'%s' object has no attribute 'conceptPhoneCounts'" % type(self).__name__)
    eng = self._eng
    cAn = self._gather_rows(eng.cA, eng.pk.phone_off)
    ph = self._gather_rows(eng.PH, eng.pk.phone_off)
    return [c[:, :, None] * p[:, None, :] for c, p in zip(cAn, ph)]

  # ------------------------------------------------------------------ I/O
  def printModel(self, fileName):
    """:685-706"""
    initFile = open(fileName+'_initialprobs.txt', 'w')
    for nState in sorted(self.lenProb):
      for i in range(nState):
        initFile.write('%d\t%d\t%f\n' % (nState, i, self.init[nState][i]))
    initFile.close()
    transFile = open(fileName+'_transitionprobs.txt', 'w')
    for nState in sorted(self.lenProb):
      for i in range(nState):
        for j in range(nState):
          transFile.write('%d\t%d\t%d\t%f\n' % (nState, i, j, self.trans[nState][i][j]))
    transFile.close()
    np.save(fileName+'_phoneprobs.npy', self.phoneProbs)
    with open(fileName+'_phone2idx.json', 'w') as f:
      json.dump(self.phone2idx, f)
    np.save(fileName+'_visualanchors.npy', self.musV)
    np.save(fileName+'_audioanchors.npy', self.musA)

  def printAlignment(self, filePrefix, isPhoneme=True, debug=False, _zero_concept_alignment=False):
    """:708-740 -- batched Viterbi / cluster / argmax launches instead of the per-pair loop.
    phone_clusters = argmax_ph sum_k conceptPhoneCounts[t] (= argmax of the frame posterior up to the
    positive per-frame factor), concept_alignment = argmax_k sum_ph conceptPhoneCounts[t]."""
    if not getattr(self, '_cA_valid', False):                  # self.conceptPhoneCounts missing (:716)
      raise AttributeError("'%s' object has no attribute 'conceptPhoneCounts'" % type(self).__name__)
    eng = self._push()
    torch = eng.torch
    pk = eng.pk
    # the argmax inputs are the LAST E-step's buffers (the reference reads the stored counts)
    ca = torch.empty((max(pk.n_phones_total, 1),), dtype=torch.int32, device=eng.device)
    from .. import _lib
    from ..engine import _ptr
    _lib.check(eng.lib.mwd_argmax_rows(_ptr(eng.cA), pk.n_phones_total, eng.K, _ptr(ca), eng._stream()))
    cas = self._gather_rows(ca[:pk.n_phones_total], pk.phone_off)
    pcs = self._gather_rows(eng.phone_clusters[:pk.n_phones_total], pk.phone_off)
    cps = self._gather_rows(eng.cC, pk.region_off)
    ali, ic, ap = eng.decode(floor_norm=True, want_probs=True, width=self._width())
    alis = self._gather_rows(ali, pk.phone_off)
    ics = self._gather_rows(ic, pk.region_off)
    aps = self._gather_rows(ap, pk.ap_offsets())
    rank, _ = self._dist()
    if rank != 0:
      return
    f = open(filePrefix+'.txt', 'w')
    aligns = []
    for i in range(len(self.vCorpus)):
      n = len(ics[i])
      aligns.append({
            'index': i,
            'image_concepts': [int(c) for c in ics[i]],
            'phone_clusters': [int(c) for c in pcs[i]],
            'concept_alignment': [int(c) for c in cas[i]],
            'alignment': [int(a) for a in alis[i]],
            'align_probs': np.asarray(aps[i]).reshape(-1, n).tolist(),
            'concept_probs': np.asarray(cps[i]).tolist(),
            'is_phoneme': isPhoneme
          })
      for a in alis[i]:
        f.write('%d ' % a)
      f.write('\n\n')
    f.close()
    with open(filePrefix+'.json', 'w') as f:
      json.dump(aligns, f, indent=4, sort_keys=True)

  def printUnimodalCluster(self, filePrefix):
    """:742-757 (the .txt file is opened and left empty, as in the reference)"""
    f = open(filePrefix+'.txt', 'w')
    cluster_infos = []
    eng = self._push()
    eng.posterior(self._width())
    pzs = self._gather_rows(eng.pz, eng.pk.region_off)
    for i in range(len(self.vCorpus)):
      clusterProbs = np.asarray(pzs[i])
      cluster_infos.append({
          'index': i,
          'image_concepts': np.argmax(clusterProbs, axis=1).tolist(),
          'cluster_probs': clusterProbs.tolist()
        })
    f.close()
    with open(filePrefix+'.json', 'w') as f:
      json.dump(cluster_infos, f, indent=4, sort_keys=True)

  def simulatedAnnealing(self, numIterations=100, T0=0.5, stepScale=5., debug=False):
    """:153-193"""
    self.trainUsingEM(numIterations=5, warmStart=False, printStatus=True)
    E0 = -self.computeAvgLogLikelihood()
    Emin = E0
    count = 0
    for epoch in range(numIterations):
      print('Simulated Annealing Iteration %d' % epoch)
      begin_time = time.time()
      init_prev = deepcopy(self.init)
      trans_prev = deepcopy(self.trans)
      phoneProbs_prev = deepcopy(self.phoneProbs)
      musV_prev = deepcopy(self.musV)
      musA_prev = deepcopy(self.musA)
      self.musV += stepScale * np.random.normal(size=(self.nWords, self.imageFeatDim))
      self.musA += stepScale * np.random.normal(size=(self.nPhones, self.audioFeatDim))
      self.trainUsingEM(numIterations=5, warmStart=True, printStatus=False)
      E1 = -self.computeAvgLogLikelihood()
      print('Current and previous energy level: ', E1, E0)
      Tk = T0 / np.log(epoch+2)
      if E1 > E0 and random.random() > np.exp(-(E1 - E0) / Tk):
        self.musV = musV_prev
        self.musA = musA_prev
        self.init = init_prev
        self.trans = trans_prev
        self.phoneProbs = phoneProbs_prev
      else:
        if debug:
          print('Random jump at temperature %.5f' % Tk)
        E0 = E1
        if E1 < Emin:
          Emin = E1
          count += 1
          print('Update %d after %.2f s: current lowest energy level is %.5f' % (count, time.time()-begin_time, Emin))
          self.printModel(self.modelName+'_%d' % count)
          self.printAlignment(self.modelName+'_%d_alignment' % count, debug=False)
          begin_time = time.time()
