"""Drop-in for hmm/audio_hmm_word_discoverer.py (``AudioHMMWordDiscoverer``, log-domain HMM aligner
with a NULL state; hmm/hmm_word_discoverer_logscale.py is the same class).  The reference's quirks
are kept: un-normalised init counts, transition counts from the last time step only, counts that
accumulate over epochs, log-likelihood printed after the M-step.  CUDA hot path (mwd_hmm_*)."""
import numpy as np
import math
import json
import random
from scipy.special import logsumexp

from ._plain_base import PlainHMMBase
from ..engine_hmm import SegmentHMMEngine

NULL = "NULL"
DEBUG = False


class AudioHMMWordDiscoverer(PlainHMMBase):
  LOG = True

  def __init__(self, *args, **kwargs):
    """Two call forms, both used by the reference code base:
      * AudioHMMWordDiscoverer(trainingCorpusFile, initProbFile=None, transProbFile=None,
        obsProbFile=None, modelName="audio_hmm_word_discoverer")
        -- the shipped discrete class (audio_hmm_word_discoverer.py:13);
      * AudioHMMWordDiscoverer(numMixtures, frameDim, fCorpus=<list of (S x embedDim) arrays>,
        tCorpus=<list of [NULL]+tokens>, initProbFile=, transProbFile=, obsModelFile=, initMethod=,
        maxLen=, fixedVariance=) -- the form SegEmbedHMMWordDiscoverer instantiates
        (audio_segembed_hmm_word_discoverer.py:86-92, the commented constructor :15): the same
        log-domain HMM with diagonal-Gaussian(-mixture) emissions per concept word."""
    file_form = ('trainingCorpusFile' in kwargs) or (len(args) > 0 and isinstance(args[0], str))
    if file_form:
      names = ['trainingCorpusFile', 'initProbFile', 'transProbFile', 'obsProbFile', 'modelName']
    else:
      names = ['numMixtures', 'frameDim']
    if len(args) > len(names):
      raise TypeError('__init__() takes at most %d positional arguments' % len(names))
    kw = dict(zip(names, args))
    for k, v in kwargs.items():
      if k in kw:
        raise TypeError("__init__() got multiple values for argument '%s'" % k)
      kw[k] = v
    trainingCorpusFile = kw.get('trainingCorpusFile', kw.get('numMixtures'))
    initProbFile, transProbFile = kw.get('initProbFile'), kw.get('transProbFile')
    obsProbFile = kw.get('obsProbFile')
    modelName = kw.get('modelName', "audio_hmm_word_discoverer")
    fCorpus, tCorpus = kw.get('fCorpus'), kw.get('tCorpus')
    obsModelFile, initMethod = kw.get('obsModelFile'), kw.get('initMethod', "rand")
    maxLen, fixedVariance, device = kw.get('maxLen', 2000), kw.get('fixedVariance', 0.02), kw.get('device')
    self.modelName = modelName
    self.fCorpus = []
    self.tCorpus = []
    self.init = {}
    self._obs_dict = {}
    self._obs_dense = None
    self.trans = {}
    self.lenProb = {}
    self.avgLogTransProb = float('-inf')
    self._device = device
    self.continuous = not isinstance(trainingCorpusFile, str)
    if self.continuous:
      self.numMixtures = int(kw['numMixtures'])
      self.frameDim = kw.get('frameDim')
      self.fCorpus = [np.asarray(x)[:maxLen] for x in fCorpus]
      self.tCorpus = [list(t) for t in tCorpus]
      self.fixedVariance = fixedVariance
      self.obsModelFile = obsModelFile
      self.initProbFile = initProbFile
      self.transProbFile = transProbFile
      self.obsProbFile = None
      self.computeTranslationLengthProbabilities()
      for m in self.lenProb:
        self.init[m] = np.log(1. / m) * np.ones((m,))
      for m in self.lenProb:
        self.trans[m] = np.log(1. / m) * np.ones((m, m))
      self._load_param_files()
      self._init_emission_model(initMethod)
    else:
      self.initialize(trainingCorpusFile)
      self.initProbFile = initProbFile
      self.transProbFile = transProbFile
      self.obsProbFile = obsProbFile
    print("Finish initialization of obs model")

  # ------------------------------------------------------------------ continuous (segment) form
  def _init_emission_model(self, initMethod):
    tv, _ = self._vocab_cont()
    Vt, M, D = len(tv), self.numMixtures, self.fCorpus[0].shape[1]
    self.featDim = D
    self.mixturePriors = np.log(np.ones((Vt, M)) / M)
    self.transVars = (self.fixedVariance if self.fixedVariance > 0 else 1.) * np.ones((Vt, M, D))
    self.transMeans = np.zeros((Vt, M, D))
    if self.obsModelFile:
      with open(self.obsModelFile + '_mixture_priors.json') as f:
        pri = json.load(f)
      with open(self.obsModelFile + '_translation_means.json') as f:
        mu = json.load(f)
      with open(self.obsModelFile + '_translation_variances.json') as f:
        vr = json.load(f)
      for w, i in tv.items():
        if w in mu:
          self.mixturePriors[i] = np.array(pri[w])
          self.transMeans[i] = np.array(mu[w])
          self.transVars[i] = np.array(vr[w])
      return
    # initMethod "rand": every mixture mean starts at a random segment of a caption containing the word
    occ = {}
    for ex, ts in enumerate(self.tCorpus):
      for w in set(ts):
        occ.setdefault(w, []).append(ex)
    for w, i in tv.items():
      for m in range(M):
        ex = occ[w][np.random.randint(len(occ[w]))]
        self.transMeans[i, m] = self.fCorpus[ex][np.random.randint(len(self.fCorpus[ex]))]

  def _vocab_cont(self):
    if getattr(self, '_tv', None) is None:
      self._tv = {w: i for i, w in enumerate(sorted({w for e in self.tCorpus for w in e}))}
      self._tw = sorted(self._tv)
      self._tgt_ids = [np.array([self._tv[w] for w in e], dtype=np.int32) for e in self.tCorpus]
      self._fv = {}
    return self._tv, self._fv

  def _seg_engine(self):
    if getattr(self, '_seng', None) is None:
      rank, world = self._dist()
      self._vocab_cont()
      self._seng = SegmentHMMEngine(self._tgt_ids, self.fCorpus, len(self._tv), self.numMixtures,
                                    device=self._device, rank=rank, world=world)
    eng = self._seng
    eng.set_chain_params(self.init, self.trans)
    eng.set_emission_params(self.mixturePriors, self.transMeans, self.transVars)
    return eng

  def _pull_seg(self, eng):
    init, trans, lprior, means, var = eng.get_all_params()
    for m in init:
      self.init[m] = init[m]
      self.trans[m] = trans[m]
    self.mixturePriors, self.transMeans, self.transVars = lprior, means, var

  def initialize(self, fileName):
    """reference :41-70"""
    self._read_blocks(fileName, add_null=True)
    self.computeTranslationLengthProbabilities()
    for m in self.lenProb:
      self.init[m] = np.log(1. / m) * np.ones((m,))
    for m in self.lenProb:
      self.trans[m] = np.log(1. / m) * np.ones((m, m))

  def initializeModel(self):
    """reference :100-138 -- every co-occurring (tw, fw) counts once, row-normalised, log"""
    obs = self._load_param_files()
    if obs is not None:
      self.obs = obs
      return
    tv, fv = self._vocab()
    seen = np.zeros((len(tv), len(fv)), dtype=bool)
    for e, f in zip(self._tgt_ids, self._src_ids):
      seen[np.ix_(np.unique(e), np.unique(f))] = True
    dense = np.full(seen.shape, np.nan)
    cnt = seen.sum(1, keepdims=True).astype(float)
    with np.errstate(divide='ignore', invalid='ignore'):
      val = np.log(1.0 / cnt) * np.ones((1, len(fv)))
    dense[seen] = val[seen]
    self._obs_dense = dense
    self._obs_dict = None

  def trainUsingEM(self, numIterations=30, writeModel=False):
    """reference :316-393"""
    if self.continuous:
      return self._train_continuous(numIterations, writeModel)
    if writeModel:
      self.printModel('initial_model.txt')
    self.initializeModel()
    if min(len(f) for f in self.fCorpus) < 2:
      raise NameError("name 'transJumpCount' is not defined")   # the reference fails the same way (:223)
    eng = self._push()
    eng.reset_accumulators()
    N = len(self.tCorpus)
    for epoch in range(numIterations):
      eng.em_iteration()
      print('Epoch', epoch, 'Average Log Likelihood:', float(eng.loglik_sum()) / N)
      if writeModel:
        self._pull(eng)
        self.printModel(self.modelName + 'model_iter=' + str(epoch))
    self._pull(eng)

  def printAlignment(self, filePrefix, isPhoneme=True):
    """reference :454-484"""
    self._print_alignment(filePrefix, lambda fSen: {'is_phoneme': False, 'is_audio': True})

  def _train_continuous(self, numIterations, writeModel):
    if writeModel:
      self.printModel('initial_model.txt')
    if min(len(f) for f in self.fCorpus) < 2:
      raise NameError("name 'transJumpCount' is not defined")
    eng = self._seg_engine()
    eng.reset_accumulators()
    N = len(self.tCorpus)
    for epoch in range(numIterations):
      eng.em_iteration(update_var=(self.fixedVariance <= 0))
      print('Epoch', epoch, 'Average Log Likelihood:', float(eng.loglik_sum()) / N)
      if writeModel:
        self._pull_seg(eng)
        self.printModel(self.modelName + 'model_iter=' + str(epoch))
    self._pull_seg(eng)

  def computeAvgLogLikelihood(self):
    if self.continuous:
      return float(self._seg_engine().loglik_sum()) / len(self.tCorpus)
    return PlainHMMBase.computeAvgLogLikelihood(self)

  def align(self, fSen, eSen, unkProb=10e-12):
    """discrete form: reference :396-427.  continuous form: align(embeddings, tSen) ->
    (assignment per segment, per-segment log scores) as SegEmbedHMMWordDiscoverer.assign expects
    (audio_segembed_hmm_word_discoverer.py:167-170)."""
    if not self.continuous:
      return PlainHMMBase.align(self, fSen, eSen, unkProb)
    tv, _ = self._vocab_cont()
    e = np.array([tv[w] for w in eSen], dtype=np.int32)
    x = np.asarray(fSen)
    main = self._seg_engine()
    mini = SegmentHMMEngine([e], [x], len(tv), self.numMixtures, device=self._device)
    for dst, src in ((mini.init_t, main.init_t), (mini.trans_t, main.trans_t), (mini.lprior, main.lprior),
                     (mini.means, main.means), (mini.var, main.var)):
      dst.copy_(src)
    ali, ap = mini.align(unkProb)
    n, T = len(e), len(x)
    return [int(a) for a in ali.cpu().numpy()], ap.cpu().numpy().reshape(T - 1, n).tolist()

  def printModel(self, fileName):
    if not self.continuous:
      return PlainHMMBase.printModel(self, fileName)
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
    tv, _ = self._vocab_cont()
    for name, arr in (('mixture_priors', self.mixturePriors), ('translation_means', self.transMeans),
                      ('translation_variances', self.transVars)):
      with open(fileName + '_obs_model_' + name + '.json', 'w') as f:
        json.dump({w: np.asarray(arr[i]).tolist() for w, i in tv.items()}, f)
