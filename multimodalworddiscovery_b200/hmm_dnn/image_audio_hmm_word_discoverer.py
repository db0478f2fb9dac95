"""Drop-in for hmm_dnn/image_audio_hmm_word_discoverer.py (``ImageAudioHMMWordDiscoverer``,
SURVEY 8 f2; driver: run_image2audio.py:232-259).

Same module-level names, constructor and method signatures, files in and out as the reference
module; the EM hot path is CUDA (``engine_audio.IKAudioEngine``): frame posterior GEMM, dense
emission, the (region, concept)-state recursion kernels shared with the image-phone classes and
the concept-phone count reduction.

Reference quirks kept on purpose:
  * only the first 30 pairs of both feature files are read (:63, :80);
  * ``updateSoftmaxWeightA``'s gradient vanishes identically (see engine_audio.mstep), so
    ``WA`` only decays by ``1 - momentum``;
  * ``simulatedAnnealing`` reads ``self.W``, which this class never defines (:173) -> AttributeError
    after the initial 5 EM iterations, as in the reference.
"""
import numpy as np
import math
import json
import time
from scipy.special import logsumexp
import random
from copy import deepcopy

from ._ik_base import ImagePhoneHMMBase, write_alignment_files

NULL = "NULL"
DEBUG = False
EPS = 1e-50
random.seed(1)
np.random.seed(1)


class ImageAudioHMMWordDiscoverer(ImagePhoneHMMBase):
  GAUSSIAN = False

  def __init__(self, speechFeatureFile, imageFeatureFile, modelConfigs, modelName='image_phone_hmm_word_discoverer'):
    self.modelName = modelName
    self.aCorpus = []
    self.vCorpus = []
    self.hasNull = modelConfigs.get('has_null', False)
    self.nWords = modelConfigs.get('n_words', 66)
    self.nPhones = modelConfigs.get('n_phones', 42)
    self.momentum = modelConfigs.get('momentum', 0.)
    self.lr = modelConfigs.get('learning_rate', 10.)
    self.normalize_vfeat = modelConfigs.get('normalize_vfeat', False)
    self.initProbFile = modelConfigs.get('init_prob_file', None)
    self.transProbFile = modelConfigs.get('trans_prob_file', None)
    self.phoneProbFile = modelConfigs.get('phone_prob_file', None)
    self.audioPosteriorFile = modelConfigs.get('audio_posterior_weights_file', None)
    self.imagePosteriorFile = modelConfigs.get('image_posterior_weights_file', None)
    # optional, B200-build-only keys
    self._device = modelConfigs.get('device', None)
    self._feature_dtype = modelConfigs.get('feature_dtype', 'auto')
    self._pair_limit = modelConfigs.get('pair_limit', 30)     # the reference's hard-wired [:30]

    self.init = {}
    self.trans = {}
    self.lenProb = {}
    self.phoneProbs = None
    self.avgLogTransProb = float('-inf')
    self.readCorpus(speechFeatureFile, imageFeatureFile, debug=False)

  def readCorpus(self, speechFeatFile, imageFeatFile, debug=False):
    """reference :43-91"""
    self.phone2idx = {}
    vCorpus = self._read_features(imageFeatFile)
    if self.normalize_vfeat:
      vCorpus = [(vSen.T / np.linalg.norm(vSen, ord=2, axis=-1)).T for vSen in vCorpus]
    if self.hasNull:
      # the reference reads self.imageFeatDim before assigning it (:59) and raises here
      vCorpus = [np.concatenate((np.zeros((1, self.imageFeatDim)), vfeat), axis=0) for vfeat in vCorpus]
    self.vCorpus = vCorpus[:self._pair_limit]                  # :63
    self.imageFeatDim = self.vCorpus[0].shape[-1]
    nImages = 0
    for ex, vfeat in enumerate(self.vCorpus):
      nImages += len(vfeat)
      if vfeat.shape[-1] == 0:
        print('example {} is empty:'.format(ex), vfeat.shape)
        self.vCorpus[ex] = np.zeros((1, self.imageFeatDim))
    self.aCorpus = self._read_features(speechFeatFile, limit=self._pair_limit)   # :80
    nTokens = 0
    self.audioFeatDim = self.aCorpus[0].shape[-1]
    for afeat in self.aCorpus:
      nTokens += afeat.shape[0]
    print('----- Corpus Summary -----')
    print('Number of examples: ', len(self.aCorpus))
    print('Number of phonetic categories: ', self.nPhones)
    print('Number of phones: ', nTokens)
    print('Number of objects: ', nImages)
    print("Number of word clusters: ", self.nWords)

  def initializeModel(self, alignments=None):
    """reference :93-145"""
    begin_time = time.time()
    self.computeTranslationLengthProbabilities()
    for m in self.lenProb:
      self.init[m] = 1. / m * np.ones((m,))
    for m in self.lenProb:
      self.trans[m] = 1. / m * np.ones((m, m))
    self._load_init_trans_files(create_missing=False)
    if self.phoneProbFile:
      self.phoneProbs = np.load(self.phoneProbFile)
    else:
      self.phoneProbs = 1. / self.nPhones * np.ones((self.nWords, self.nPhones))
    if self.imagePosteriorFile:
      imagePosteriorWeights = np.load(self.imagePosteriorFile)
      weight_v, bias_v = imagePosteriorWeights['weight'], imagePosteriorWeights['bias']
      self.WV = np.concatenate([weight_v, bias_v[:, np.newaxis]], axis=1)
    else:
      self.WV = .1 * np.random.normal(size=(self.nWords, self.imageFeatDim + 1))
      self.WV[:, -1] = 0.
    if self.audioPosteriorFile:
      audioPosteriorWeights = np.load(self.audioPosteriorFile)
      weight_a, bias_a = audioPosteriorWeights['weight'], audioPosteriorWeights['bias']
      self.WA = np.concatenate([weight_a, bias_a[:, np.newaxis]], axis=1)
    else:
      self.WA = .1 * np.random.normal(size=(self.nPhones, self.audioFeatDim + 1))
      self.WA[:, -1] = 0.
    print("Finish initialization after %0.3f s" % (time.time() - begin_time))

  # ------------------------------------------------------------------ engine plumbing
  def _phone_ids(self):
    # only the lengths matter (computeTranslationLengthProbabilities)
    return [np.zeros(len(a), dtype=np.int32) for a in self.aCorpus]

  def _engine(self):
    from ..engine_audio import IKAudioEngine, pack_audio_pairs
    token = (id(self.vCorpus), len(self.vCorpus), id(self.aCorpus), len(self.aCorpus), self.nWords, self.nPhones)
    if getattr(self, '_eng', None) is None or self._eng_token != token:
      rank, world = self._dist()
      from ..corpus import resolve_feature_dtype
      dt = resolve_feature_dtype(self._feature_dtype, self.vCorpus, self.aCorpus)
      pk, audio = pack_audio_pairs(self.vCorpus, self.aCorpus, feat_dtype=dt, rank=rank, world=world)
      self._eng = IKAudioEngine(pk, audio, self.nWords, self.nPhones, device=self._device)
      self._eng_token = token
      self._cA_valid = False
    return self._eng

  def _push(self):
    eng = self._engine()
    eng.set_params(self.init, self.trans, self.phoneProbs, self.WV)
    eng.set_audio_param(self.WA)
    return eng

  def _pull(self, eng):
    init, trans, pp, WV = eng.get_params()
    for m in init:
      self.init[m] = init[m]
      self.trans[m] = trans[m]
    self.phoneProbs = pp
    self.WV = WV
    self.WA = eng.get_audio_param()

  # ------------------------------------------------------------------ single-pair API
  def softmaxLayerV(self, vSen, debug=False):
    """:543-547"""
    return self._push().posterior_rows(np.asarray(vSen))

  def softmaxLayerA(self, aSen, debug=False):
    """:549-554"""
    ph, _ = self._push().emission_rows(np.asarray(aSen))
    return ph.cpu().numpy()

  def forward(self, vSen, aSen, debug=False):
    """:284-306 -> (T, n, K)"""
    return self._push().dense_sweep_audio(np.asarray(vSen), np.asarray(aSen), backward=False)

  def backward(self, vSen, aSen, debug=False):
    """:316-340 -> (T, n, K)"""
    return self._push().dense_sweep_audio(np.asarray(vSen), np.asarray(aSen), backward=True)

  def align(self, aSen, vSen, unkProb=10e-12, debug=False):
    """:604-647"""
    ali, ap, _, _ = self._push().decode_pair_audio(np.asarray(vSen), np.asarray(aSen))
    return [int(a) for a in ali], ap.tolist()

  def cluster(self, aSen, vSen, alignment):
    """:649-661"""
    _, _, ic, cs = self._push().decode_pair_audio(np.asarray(vSen), np.asarray(aSen), alignment=np.asarray(alignment))
    return [int(c) for c in ic], cs.tolist()

  # ------------------------------------------------------------------ I/O
  def printModel(self, fileName):
    """:663-684"""
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
    np.save(fileName+'_visual_posterior_weights.npy', self.WV)
    np.save(fileName+'_audio_posterior_weights.npy', self.WA)

  def printAlignment(self, filePrefix, isPhoneme=True, debug=False, _zero_concept_alignment=False):
    """:687-721 -- one batched Viterbi + cluster launch instead of the per-pair loop."""
    eng = self._push()
    ali, ic, ap = eng.decode(floor_norm=True, want_probs=True)
    pk = eng.pk
    alis = self._gather_rows(ali, pk.phone_off)
    ics = self._gather_rows(ic, pk.region_off)
    aps = self._gather_rows(ap, pk.ap_offsets())
    rank, _ = self._dist()
    if rank != 0:
      return
    write_alignment_files(filePrefix, alis, ics, aps, n_concepts=self.nWords, is_phoneme=isPhoneme)

  def simulatedAnnealing(self, numIterations=100, T0=0.5, stepScale=5., debug=False):
    """:157-192.  The reference snapshots ``self.W`` (:173), an attribute this class does not have."""
    self.trainUsingEM(numIterations=5, warmStart=False, printStatus=True)
    E0 = -self.computeAvgLogLikelihood()
    W_prev = deepcopy(self.W)       # AttributeError, exactly as in the reference
    raise AssertionError('unreachable: %r %r' % (E0, W_prev))
