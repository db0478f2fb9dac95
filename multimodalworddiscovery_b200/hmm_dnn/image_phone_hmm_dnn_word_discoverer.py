"""Drop-in for hmm_dnn/image_phone_hmm_dnn_word_discoverer.py (``ImagePhoneHMMDNNWordDiscoverer``,
the "two-layer" model of run_image2phone.py --model_type two-layer): same HMM as the linear class,
image posterior softmax(W [relu(V [v;1]); 1]), back-propagated gradient M-step, EPS-floored
init/trans, un-floored Viterbi scores.  SURVEY 8(f1).  CUDA hot path (see ``_ik_base.py``)."""
import numpy as np
import math
import json
import time
from scipy.special import logsumexp
import random
from copy import deepcopy

from ._ik_base import write_alignment_files, ImagePhoneHMMBase, OneHotCorpus, one_hot_to_ids

NULL = "NULL"
DEBUG = False
EPS = 1e-50
random.seed(1)
np.random.seed(1)


class ImagePhoneHMMDNNWordDiscoverer(ImagePhoneHMMBase):
  GAUSSIAN = False
  TWO_LAYER = True

  def __init__(self, speechFeatureFile, imageFeatureFile, modelConfigs, initProbFile=None, transProbFile=None, obsProbFile=None, modelName='image_phone_hmm_word_discoverer'):
    self.modelName = modelName
    self.aCorpus = []
    self.vCorpus = []
    self.hasNull = modelConfigs.get('has_null', False)
    self.nWords = modelConfigs.get('n_words', 66)
    self.hiddenDim = modelConfigs.get('hidden_dim', 100)
    self.momentum = modelConfigs.get('momentum', 0.)
    self.lr = modelConfigs.get('learning_rate', 10.)
    self.imagePosteriorFile = modelConfigs.get('image_posterior_weights_file', None)
    self.normalize_vfeat = modelConfigs.get('normalize_vfeat', False)
    self._device = modelConfigs.get('device', None)
    self._feature_dtype = modelConfigs.get('feature_dtype', 'auto')
    self._posterior_precision = modelConfigs.get('posterior_precision', 'float64')   # or 'mixed' (see _lib.mixed_bits)
    self._keep_cA = False                      # the reference class keeps no conceptCountsA

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
    """reference :43-104 (same as the linear class)"""
    vCorpus = self._read_features(imageFeatFile)
    if self.normalize_vfeat:
      vCorpus = [(vSen.T / np.linalg.norm(vSen, ord=2, axis=-1)).T for vSen in vCorpus]
    self.vCorpus = vCorpus
    if self.hasNull:
      self.vCorpus = [np.concatenate((np.zeros((1, self.imageFeatDim)), vfeat), axis=0) for vfeat in self.vCorpus]
    self.imageFeatDim = self.vCorpus[0].shape[-1]
    ids, nTypes, nPhones = self._read_captions(speechFeatFile)
    self._finish_corpus(ids, nTypes, nPhones)

  def initializeModel(self, alignments=None):
    """reference :107-148"""
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
      V_weight, V_bias = posteriorWeights['arr_0'], posteriorWeights['arr_1']
      W_weight, W_bias = posteriorWeights['arr_2'], posteriorWeights['arr_3']
      self.V = np.concatenate([V_weight, V_bias[:, np.newaxis]], axis=1)
      self.W = np.concatenate([W_weight, W_bias[:, np.newaxis]], axis=1)
      self.hiddenDim = self.V.shape[0]
    else:
      self.V = np.random.normal(size=(self.hiddenDim, self.imageFeatDim + 1))
      self.W = np.random.uniform(low=-1., high=1., size=(self.nWords, self.hiddenDim + 1))
    print("Finish initialization after %0.3f s" % (time.time() - begin_time))

  def trainUsingEM(self, numIterations=20,
                         writeModel=False,
                         warmStart=False,
                         convergenceEpsilon=0.01,
                         printStatus=True,
                         freezeTransition=False,
                         debug=False):
    """reference :209-290"""
    return ImagePhoneHMMBase.trainUsingEM(self, numIterations, writeModel, warmStart, convergenceEpsilon,
                                          printStatus, debug, _freeze_trans=freezeTransition)

  def hiddenLayer(self, vSen, debug=False, bias=False):
    """reference :573-579 (computed on the GPU)"""
    vSen = np.asarray(vSen)
    if len(vSen.shape) == 1:
      vSen = vSen[np.newaxis]
    eng = self._push()
    torch = eng.torch
    dt = np.float64 if eng.feat_is_f64 else np.float32
    v_d = torch.from_numpy(np.ascontiguousarray(vSen, dtype=dt)).to(eng.device)
    h = torch.empty((vSen.shape[0], eng.H), dtype=torch.float64, device=eng.device)
    from .. import _lib
    import ctypes as C
    _lib.check(eng.lib.mwd_hidden_relu(C.c_void_p(v_d.data_ptr()), eng.feat_is_f64, vSen.shape[0], eng.D,
                                       C.c_void_p(eng.V_t.data_ptr()), eng.H, C.c_void_p(h.data_ptr()), eng._stream()))
    return h.cpu().numpy()

  def softmaxLayer(self, vHidden, debug=False, bias=False):
    """reference :581-590 -- NOTE: takes the HIDDEN activations, unlike the one-layer classes"""
    vHidden = np.asarray(vHidden, dtype=np.float64)
    if len(vHidden.shape) == 1:
      vHidden = vHidden[np.newaxis]
    eng = self._push()
    torch = eng.torch
    from .. import _lib
    import ctypes as C
    h = torch.from_numpy(np.ascontiguousarray(vHidden)).to(eng.device)
    out = torch.empty((vHidden.shape[0], eng.K), dtype=torch.float64, device=eng.device)
    _lib.check(eng.lib.mwd_posterior_linear(C.c_void_p(h.data_ptr()), 1, vHidden.shape[0], eng.H,
                                            C.c_void_p(eng.post.data_ptr()), eng.K, C.c_void_p(out.data_ptr()),
                                            eng._stream()))
    return out.cpu().numpy()

  def printModel(self, fileName):
    """reference :659-679"""
    ImagePhoneHMMBase.printModel(self, fileName)
    np.save(fileName + '_softmaxweights.npy', self.W)
    np.save(fileName + '_hiddenweights.npy', self.V)

  def printAlignment(self, filePrefix, isPhoneme=True, debug=False, _zero_concept_alignment=False):
    """reference :682-716 (cluster_probs instead of concept_alignment, plus <prefix>_clusters.txt)"""
    eng = self._push()
    ali, ic, ap, cs = eng.decode(want_probs=True, want_cluster_scores=True)
    pk = eng.pk
    alis = self._gather_rows(ali, pk.phone_off)
    ics = self._gather_rows(ic, pk.region_off)
    aps = self._gather_rows(ap, pk.ap_offsets())
    css = self._gather_rows(cs, pk.region_off)
    rank, _ = self._dist()
    if rank != 0:
      return
    write_alignment_files(filePrefix, alis, ics, aps, cluster_probs=css, n_concepts=self.nWords,
                          is_phoneme=isPhoneme)
    with open(filePrefix + '_clusters.txt', 'w') as f2:
      for i in range(len(self.vCorpus)):
        f2.write(''.join('%d ' % c for c in ics[i]))
        f2.write('\n\n')
