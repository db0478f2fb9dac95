"""Drop-in for hmm/audio_hmm_word_discoverer.py (``AudioHMMWordDiscoverer``, log-domain HMM aligner
with a NULL state; hmm/hmm_word_discoverer_logscale.py is the same class).  The reference's quirks
are kept: un-normalised init counts, transition counts from the last time step only, counts that
accumulate over epochs, log-likelihood printed after the M-step.  CUDA hot path (mwd_hmm_*)."""
import numpy as np
import math
import json
from scipy.special import logsumexp

from ._plain_base import PlainHMMBase

NULL = "NULL"
DEBUG = False


class AudioHMMWordDiscoverer(PlainHMMBase):
  LOG = True

  def __init__(self, trainingCorpusFile, initProbFile=None, transProbFile=None, obsProbFile=None,
  modelName="audio_hmm_word_discoverer"):
    self.modelName = modelName
    self.fCorpus = []
    self.tCorpus = []
    self.init = {}
    self._obs_dict = {}
    self._obs_dense = None
    self.trans = {}
    self.lenProb = {}
    self.avgLogTransProb = float('-inf')
    self.initialize(trainingCorpusFile)
    self.initProbFile = initProbFile
    self.transProbFile = transProbFile
    self.obsProbFile = obsProbFile
    print("Finish initialization of obs model")

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
