"""Drop-in for hmm_dnn/image_phone_hmm_word_discoverer.py (``ImagePhoneHMMWordDiscoverer``).

Same module-level names as the reference module (drivers star-import it and rely on ``np``,
``json``, ``math``, ``time`` leaking through, run_image2phone.py:132,137), same constructor and
method signatures, same files in and out; the EM hot path is CUDA (see ``_ik_base.py``).
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


# A word discovery model using image regions and phones
# * The transition matrix is assumed to be Toeplitz
class ImagePhoneHMMWordDiscoverer(ImagePhoneHMMBase):
  GAUSSIAN = False

  def __init__(self, speechFeatureFile, imageFeatureFile, modelConfigs, initProbFile=None, transProbFile=None, obsProbFile=None, modelName='image_phone_hmm_word_discoverer'):
    self.modelName = modelName
    self.aCorpus = []
    self.vCorpus = []
    self.hasNull = modelConfigs.get('has_null', False)
    self.nWords = modelConfigs.get('n_words', 66)
    self.momentum = modelConfigs.get('momentum', 0.)
    self.lr = modelConfigs.get('learning_rate', 10.)
    self.normalize_vfeat = modelConfigs.get('normalize_vfeat', False)
    self.imagePosteriorFile = modelConfigs.get('image_posterior_weights_file', None)
    # optional, B200-build-only keys (defaults keep unchanged drivers working)
    self._device = modelConfigs.get('device', None)
    self._feature_dtype = modelConfigs.get('feature_dtype', 'auto')
    self._posterior_precision = modelConfigs.get('posterior_precision', 'float64')   # or 'mixed' (see _lib.mixed_bits)
    self._keep_cA = modelConfigs.get('keep_concept_counts_a', False)   # conceptCountsA is materialised on access

    self.init = {}
    self.trans = {}
    self.lenProb = {}
    self.obs = None
    self.avgLogTransProb = float('-inf')

    self.readCorpus(speechFeatureFile, imageFeatureFile, debug=False)
    self.initProbFile = initProbFile
    self.transProbFile = transProbFile
    self.obsProbFile = obsProbFile

  def readCorpus(self, speechFeatFile, imageFeatFile, debug=False):
    """reference :43-104"""
    vCorpus = self._read_features(imageFeatFile)
    if self.normalize_vfeat:
      vCorpus = [(vSen.T / np.linalg.norm(vSen, ord=2, axis=-1)).T for vSen in vCorpus]
    self.vCorpus = vCorpus
    if self.hasNull:
      # the reference reads self.imageFeatDim before assigning it (:63-66) and raises here
      self.vCorpus = [np.concatenate((np.zeros((1, self.imageFeatDim)), vfeat), axis=0) for vfeat in self.vCorpus]
    self.imageFeatDim = self.vCorpus[0].shape[-1]
    ids, nTypes, nPhones = self._read_captions(speechFeatFile)
    self._finish_corpus(ids, nTypes, nPhones)

  def initializeModel(self, alignments=None):
    """reference :106-147"""
    begin_time = time.time()
    self.computeTranslationLengthProbabilities()
    for m in self.lenProb:
      self.init[m] = 1. / m * np.ones((m,))
    for m in self.lenProb:
      self.trans[m] = 1. / m * np.ones((m, m))
    self._load_init_trans_files(create_missing=False)
    if self.obsProbFile:
      self.obs = np.load(self.obsProbFile)
    else:
      self.obs = 1. / self.audioFeatDim * np.ones((self.nWords, self.audioFeatDim))
    if self.imagePosteriorFile:
      posteriorWeights = np.load(self.imagePosteriorFile)
      weight, bias = posteriorWeights['weight'], posteriorWeights['bias']
      self.W = np.concatenate([weight, bias[:, np.newaxis]], axis=1)
    else:
      self.W = 1. * np.random.normal(size=(self.nWords, self.imageFeatDim + 1))
      self.W[:, -1] = 0.
    print("Finish initialization after %0.3f s" % (time.time() - begin_time))
