"""Drop-in for hmm/audio_segembed_hmm_word_discoverer.py (``SegEmbedHMMWordDiscoverer``): cuts MFCC
utterances at landmarks, resamples every segment to a fixed ``embedDim`` vector (CPU preprocessing,
scipy.signal.resample, exactly as the reference) and delegates EM / align to the acoustic model.

The reference class is broken as shipped (it uses ``time`` without importing it, and its acoustic
model constructor call does not match the shipped AudioHMMWordDiscoverer); this mirror keeps the
same surface and works with the B200 ``AudioHMMWordDiscoverer`` keyword form."""
import numpy as np
import math
import json
import time
from scipy.special import logsumexp
import scipy.signal as signal
import scipy.interpolate as interpolate

NULL = "NULL"
DEBUG = False
ORDER = 'C'


class SegEmbedHMMWordDiscoverer:
  def __init__(self, acousticModel, numMixtures, frameDim, embedDim,
               sourceCorpusFile, targetCorpusFile,
               landmarkFile,
               modelDir=None,
               minWordLen=20,
               maxWordLen=100,
               modelName='audio_segembed_hmm_word_discoverer', maxLen=2000):
    self.modelName = modelName
    self.acoustic_model = acousticModel
    self.initProbFile = None
    self.transProbFile = None
    self.obsModelFile = None
    if modelDir:
      self.initProbFile = modelDir + "model_final_initialprobs.txt"
      self.transProbFile = modelDir + "model_final_transitionprobs.txt"
      self.obsModelFile = modelDir + "model_final_obs_model"
    self.init = {}
    self.trans = {}
    self.lenProb = {}
    self.assignments = []
    self.segmentations = []
    self.embeddings = []
    self.numMixtures = numMixtures
    self.avgLogTransProb = float('-inf')
    self.embedDim = embedDim
    self.frameDim = frameDim
    self.fCorpus = []
    self.tCorpus = []
    self.initialize(landmarkFile, sourceCorpusFile, targetCorpusFile, maxLen=maxLen)

  def initialize(self, landmarkFile, fFileName, tFileName, initProbFile=None, transProbFile=None, obsModelFile=None, initMethod="rand", fixedVariance=0.02, maxLen=2000):
    """reference :52-93"""
    fp = open(tFileName)
    tCorpus = fp.read().split('\n')
    self.tCorpus = [[NULL] + tw.split() for tw in tCorpus]
    fp.close()
    fCorpus = np.load(fFileName)
    keys = sorted(fCorpus.keys(), key=lambda x: int(x.split('_')[-1]))
    self.fCorpus = [fCorpus[k] for k in keys]
    self.fCorpus = [fSen[:maxLen] for fSen in self.fCorpus]
    self.featDim = self.fCorpus[0].shape[1]
    self.data_ids = [feat_id.split('_')[-1] for feat_id in keys]
    landmarks = np.load(landmarkFile)
    for lm_id in sorted(landmarks, key=lambda x: int(x.split('_')[-1])):
      segmentation = []
      for b in landmarks[lm_id]:
        if b <= maxLen:
          segmentation.append(b)
        else:
          segmentation.append(maxLen)
          break
      self.segmentations.append(segmentation)
    for i, (fSen, segmentation) in enumerate(zip(self.fCorpus, self.segmentations)):
      self.embeddings.append(self.getSentEmbeds(fSen, segmentation, frameDim=self.frameDim))
    # zip() in the reference silently truncates to the shorter list
    n = min(len(self.embeddings), len(self.tCorpus))
    self.acoustic_model = self.acoustic_model(self.numMixtures, self.frameDim,
                        fCorpus=self.embeddings[:n], tCorpus=self.tCorpus[:n],
                        initProbFile=initProbFile,
                        transProbFile=transProbFile,
                        obsModelFile=obsModelFile,
                        initMethod=initMethod,
                        maxLen=maxLen, fixedVariance=fixedVariance)
    print("Finish initialization of acoustic model")

  def trainUsingEM(self, numIterations=30, numAMSteps=1, modelPrefix='', writeModel=False):
    """reference :95-111"""
    if writeModel:
      self.acoustic_model.printModel('initial_model.txt')
    for epoch in range(numIterations):
      print("Start training iteration " + str(epoch))
      begin_time = time.time()
      self.acoustic_model.trainUsingEM(numIterations=numAMSteps)
      print("Acoustic model training takes %0.5f s to finish" % (time.time() - begin_time))
      if writeModel:
        self.acoustic_model.printModel(modelPrefix + "model_iter=" + str(epoch))
    if writeModel:
      self.acoustic_model.printModel(modelPrefix + 'model_final')

  def embed(self, y, frameDim=None, technique="resample"):
    """reference :114-143"""
    if frameDim:
      y = y[:, :frameDim].T
    else:
      y = y.T
      frameDim = self.featDim
    n = int(self.embedDim / frameDim)
    if y.shape[0] == 1:
      y_new = np.repeat(y, n)
    if technique == "interpolate":
      x = np.arange(y.shape[1])
      f = interpolate.interp1d(x, y, kind="linear")
      x_new = np.linspace(0, y.shape[1] - 1, n)
      y_new = f(x_new).flatten(ORDER)
    elif technique == "resample":
      y_new = signal.resample(y.astype("float32"), n, axis=1).flatten(ORDER)
    elif technique == "rasanen":
      n_frames_in_multiple = int(np.floor(y.shape[1] / n)) * n
      y_new = np.mean(y[:, :n_frames_in_multiple].reshape((y.shape[0], n, -1)), axis=-1).flatten(ORDER)
    return y_new

  def getSentEmbeds(self, x, segmentation, frameDim=12):
    """reference :145-155"""
    n_words = len(segmentation) - 1
    embeddings = []
    for i_w in range(n_words):
      seg = x[segmentation[i_w]:segmentation[i_w + 1]]
      embeddings.append(self.embed(seg, frameDim=frameDim))
    return np.array(embeddings)

  def getSentDurations(self, segmentation):
    n_words = len(segmentation) - 1
    return [segmentation[i_w + 1] - segmentation[i_w] for i_w in range(n_words)]

  def assign(self, i):
    return self.acoustic_model.align(self.embeddings[i], self.tCorpus[i])

  def align(self, i):
    """reference :172-191 -- per-segment assignment expanded to frames by duration.  (The
    acoustic model's Viterbi scores start at the second segment, so the zip, as in the
    reference, drops the last segment's frames from align_probs but not from the alignment.)"""
    durations = self.getSentDurations(self.segmentations[i])
    assignment, assign_scores = self.assign(i)
    alignment = []
    align_probs = []
    for j, dur in zip(assignment, durations):
      alignment.extend([j] * int(dur))
    for scores, dur in zip(assign_scores, durations):
      align_probs.extend([scores] * int(dur))
    return alignment, align_probs

  def printAlignment(self, filePrefix, isPhoneme=True):
    """reference :193-222"""
    f = open(filePrefix + '.txt', 'w')
    aligns = []
    n = min(len(self.embeddings), len(self.tCorpus))
    for i, (fSen, tSen) in enumerate(zip(self.fCorpus[:n], self.tCorpus[:n])):
      alignment, alignProbs = self.align(i)
      align_info = {
            'index': self.data_ids[i],
            'image_concepts': tSen,
            'alignment': alignment,
            'align_probs': alignProbs,
            'is_phoneme': False,
            'is_audio': True
          }
      aligns.append(align_info)
      f.write('%s\n%s\n' % (tSen, fSen))
      for a in alignment:
        f.write('%d ' % a)
      f.write('\n\n')
    f.close()
    with open(filePrefix + '.json', 'w') as f:
      json.dump(aligns, f, indent=4, sort_keys=True)
