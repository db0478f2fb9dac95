"""Drop-in for hmm/hmm_word_discoverer.py (``HMMWordDiscoverer``, Vogel-96 style HMM aligner in the
probability domain).  Same surface as the reference; the EM hot path is CUDA (mwd_hmm_*)."""
import numpy as np
import math
import json

from ._plain_base import PlainHMMBase

NULL = "NULL"
DEBUG = False


class HMMWordDiscoverer(PlainHMMBase):
  LOG = False

  def __init__(self, trainingCorpusFile, initProbFile=None, transProbFile=None, obsProbFile=None, modelName='hmm_word_discoverer'):
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

  def initialize(self, fileName):
    """reference :34-64"""
    self._read_blocks(fileName, add_null=False)
    self.computeTranslationLengthProbabilities()
    for m in self.lenProb:
      self.init[m] = 1. / m * np.ones((m,))
    for m in self.lenProb:
      self.trans[m] = 1. / m * np.ones((m, m))

  def initializeModel(self):
    """reference :67-108 -- co-occurrence counts (one per token pair), row-normalised"""
    obs = self._load_param_files()
    if obs is not None:
      self.obs = obs
      return
    tv, fv = self._vocab()
    c = np.zeros((len(tv), len(fv)))
    for e, f in zip(self._tgt_ids, self._src_ids):
      c += np.outer(np.bincount(e, minlength=len(tv)), np.bincount(f, minlength=len(fv)))
    dense = np.full(c.shape, np.nan)
    with np.errstate(invalid='ignore', divide='ignore'):
      norm = c / c.sum(1, keepdims=True)
    dense[c > 0] = norm[c > 0]
    self._obs_dense = dense
    self._obs_dict = None

  def trainUsingEM(self, numIterations=80, writeModel=False, convergenceEpsilon=0.01):
    """reference :248-299 (the log-likelihood printed per epoch belongs to the entering parameters)"""
    if writeModel:
      self.printModel('initial_model.txt')
    self.initializeModel()
    eng = self._push()
    N = len(self.tCorpus)
    for epoch in range(numIterations):
      ll = eng.em_iteration()
      print('Epoch', epoch, 'Average Log Likelihood:', float(ll) / N)
      if writeModel:
        self._pull(eng)
        self.printModel(self.modelName + '_iter=' + str(epoch) + '.txt')
    self._pull(eng)

  def printAlignment(self, filePrefix, isPhoneme=True):
    """reference :355-383"""
    self._print_alignment(filePrefix, lambda fSen: {'caption': fSen, 'is_phoneme': isPhoneme})
