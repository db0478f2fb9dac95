"""Drop-in for hmm_dnn/image_phone_gaussian_hmm_word_discoverer.py
(``ImagePhoneGaussianHMMWordDiscoverer``): RBF image posterior softmax(-|v - mu_k|^2 / width),
EPS-floored init/trans M-step, anchor update.  Same surface as the reference; CUDA hot path.
"""
import numpy as np
import math
import json
import time
from scipy.special import logsumexp
import random
from copy import deepcopy

from ._ik_base import ImagePhoneHMMBase, OneHotCorpus, one_hot_to_ids

NULL = "NULL"
DEBUG = False
EPS = 1e-50
random.seed(1)
np.random.seed(1)


class ImagePhoneGaussianHMMWordDiscoverer(ImagePhoneHMMBase):
  GAUSSIAN = True

  def __init__(self, speechFeatureFile, imageFeatureFile, modelConfigs, modelName='image_phone_hmm_word_discoverer'):
    self.modelName = modelName
    self.aCorpus = []
    self.vCorpus = []
    self.hasNull = modelConfigs.get('has_null', False)
    self.nWords = modelConfigs.get('n_words', 66)
    self.width = modelConfigs.get('width', 1.)
    self.momentum = modelConfigs.get('momentum', 0.)
    self.lr = modelConfigs.get('learning_rate', 10.)
    self.isExact = modelConfigs.get('is_exact', False)
    self.normalize_vfeat = modelConfigs.get('normalize_vfeat', False)   # read but ignored, as in the reference
    self._device = modelConfigs.get('device', None)
    self._feature_dtype = modelConfigs.get('feature_dtype', 'auto')
    self._posterior_precision = modelConfigs.get('posterior_precision', 'float64')   # or 'mixed' (see _lib.mixed_bits)
    self._keep_cA = modelConfigs.get('keep_concept_counts_a', False)   # conceptCountsA is materialised on access
    # The reference silently keeps only the first 30 pairs (debug leftover, :56,:87).  Default is
    # bug-compatible; set modelConfigs['max_pairs']=None to train on the whole corpus.
    self._max_pairs = modelConfigs.get('max_pairs', 30)

    self.init = {}
    self.trans = {}
    self.lenProb = {}
    self.obs = None
    self.avgLogTransProb = float('-inf')

    self.readCorpus(speechFeatureFile, imageFeatureFile, debug=False)
    self.initProbFile = modelConfigs.get('init_prob_file', None)
    self.transProbFile = modelConfigs.get('trans_prob_file', None)
    self.obsProbFile = modelConfigs.get('obs_prob_file', None)
    self.visualAnchorFile = modelConfigs.get('visual_anchor_file', None)

  def readCorpus(self, speechFeatFile, imageFeatFile, debug=False):
    """reference :46-100"""
    self.vCorpus = self._read_features(imageFeatFile, limit=self._max_pairs)
    if self.hasNull:
      self.vCorpus = [np.concatenate((np.zeros((1, self.imageFeatDim)), vfeat), axis=0) for vfeat in self.vCorpus]
    self.imageFeatDim = self.vCorpus[0].shape[-1]
    ids, nTypes, nPhones = self._read_captions(speechFeatFile, limit=self._max_pairs)
    self._finish_corpus(ids, nTypes, nPhones)

  def initializeModel(self, alignments=None):
    """reference :102-145"""
    begin_time = time.time()
    self.computeTranslationLengthProbabilities()
    for m in self.lenProb:
      self.init[m] = 1. / m * np.ones((m,))
    for m in self.lenProb:
      self.trans[m] = 1. / m * np.ones((m, m))
    self._load_init_trans_files(create_missing=True)
    if self.obsProbFile:
      self.obs = np.load(self.obsProbFile)
    else:
      self.obs = 1. / self.audioFeatDim * np.ones((self.nWords, self.audioFeatDim))
    if self.visualAnchorFile:
      self.mus = np.load(self.visualAnchorFile)
    else:
      from sklearn.cluster import KMeans
      self.mus = KMeans(n_clusters=self.nWords).fit(np.concatenate(self.vCorpus, axis=0)).cluster_centers_
    print("Finish initialization after %0.3f s" % (time.time() - begin_time))
    self.printUnimodalCluster(filePrefix=self.modelName)

  def trainUsingEM(self, numIterations=20, writeModel=False, warmStart=False, convergenceEpsilon=0.01,
                   printStatus=True, debug=False, **kw):
    if self.isExact and numIterations > 0:
      # reference :480-486: the closed-form anchor update reads conceptCount / zProb before assigning
      # them, so the first M-step raises; there is no exact update to mirror
      raise NameError("name 'conceptCount' is not defined")
    ImagePhoneHMMBase.trainUsingEM(self, numIterations, writeModel, warmStart, convergenceEpsilon, printStatus,
                                   debug, **kw)

  def printUnimodalCluster(self, filePrefix):
    """reference :663-678"""
    f = open(filePrefix + '.txt', 'w')
    cluster_infos = []
    eng = self._push()
    eng.posterior(self._width())
    pzs = self._gather_rows(eng.pz, eng.pk.region_off)
    for i in range(len(self.vCorpus)):
      clusterProbs = pzs[i]
      clusters = np.argmax(clusterProbs, axis=1)
      cluster_infos.append({
          'index': i,
          'image_concepts': clusters.tolist(),
          'cluster_probs': clusterProbs.tolist()
        })
    f.close()
    with open(filePrefix + '.json', 'w') as f:
      json.dump(cluster_infos, f, indent=4, sort_keys=True)
