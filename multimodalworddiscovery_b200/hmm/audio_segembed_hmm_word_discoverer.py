"""Drop-in for hmm/audio_segembed_hmm_word_discoverer.py (``SegEmbedHMMWordDiscoverer``).

What the reference class does: cut every MFCC utterance at its landmark frames, turn each
variable-length segment into one fixed ``embedDim`` vector (``embedDim / frameDim`` frames obtained by FFT
resampling along time, float32 input, C-order flattening), hand the per-utterance embedding matrices
and the ``[NULL] + concepts`` captions to an acoustic HMM, and expand the HMM's per-segment decisions
back to frames for the alignment dump.

This module keeps that surface (constructor signature, ``trainUsingEM / embed / getSentEmbeds / assign /
align / printAlignment``, attributes ``embeddings / segmentations / tCorpus / fCorpus / data_ids /
acoustic_model``) but is organised around whole-utterance array operations: landmarks are clipped with
one vectorised scan, all segments of an utterance that share a length are
resampled in ONE batched ``scipy.signal.resample`` call (same FFT, so the values are those of the
reference's per-segment calls), and frame expansion is ``np.repeat``.  The reference file cannot run as
shipped (``time`` is never imported and its acoustic-model call does not fit the shipped
``AudioHMMWordDiscoverer``); here the acoustic model is the CUDA-backed one in its keyword form.
"""
import json
import math
import time

import numpy as np
import scipy.interpolate as interpolate
import scipy.signal as signal
from scipy.special import logsumexp

NULL = "NULL"
DEBUG = False
ORDER = 'C'


def _key_index(name):
    return int(name.split('_')[-1])


def _clip_landmarks(bounds, max_len):
    """Frame boundaries of one utterance, cut at ``max_len``: boundaries past it collapse into a single
    closing boundary at ``max_len`` (reference :68-76)."""
    bounds = np.asarray(bounds).ravel()
    past = np.flatnonzero(bounds > max_len)
    if past.size:
        bounds = np.concatenate([bounds[:past[0]], [max_len]])
    return [b.item() if hasattr(b, 'item') else b for b in bounds]


def _resample_rows(block, n_out):
    """FFT resampling of a (segments, frameDim, frames) float32 block to ``n_out`` frames."""
    return signal.resample(block, n_out, axis=-1)


class SegEmbedHMMWordDiscoverer:
    def __init__(self, acousticModel, numMixtures, frameDim, embedDim, sourceCorpusFile, targetCorpusFile,
                 landmarkFile, modelDir=None, minWordLen=20, maxWordLen=100,
                 modelName='audio_segembed_hmm_word_discoverer', maxLen=2000):
        self.modelName = modelName
        self.acoustic_model = acousticModel
        self.numMixtures, self.frameDim, self.embedDim = numMixtures, frameDim, embedDim
        files = ('initialprobs.txt', 'transitionprobs.txt', 'obs_model')
        self.initProbFile, self.transProbFile, self.obsModelFile = (
            tuple(modelDir + 'model_final_' + f for f in files) if modelDir else (None, None, None))
        self.init, self.trans, self.lenProb = {}, {}, {}
        self.assignments, self.segmentations, self.embeddings = [], [], []
        self.fCorpus, self.tCorpus = [], []
        self.avgLogTransProb = float('-inf')
        self.initialize(landmarkFile, sourceCorpusFile, targetCorpusFile, maxLen=maxLen)

    # ------------------------------------------------------------------ corpus -> embeddings
    def initialize(self, landmarkFile, fFileName, tFileName, initProbFile=None, transProbFile=None,
                   obsModelFile=None, initMethod="rand", fixedVariance=0.02, maxLen=2000):
        with open(tFileName) as fp:
            self.tCorpus = [[NULL] + line.split() for line in fp.read().split('\n')]
        archive = np.load(fFileName)
        names = sorted(archive.keys(), key=_key_index)
        self.data_ids = [name.split('_')[-1] for name in names]
        self.fCorpus = [archive[name][:maxLen] for name in names]
        self.featDim = self.fCorpus[0].shape[1]
        marks = np.load(landmarkFile)
        self.segmentations = [_clip_landmarks(marks[name], maxLen) for name in sorted(marks, key=_key_index)]
        self.embeddings = [self.getSentEmbeds(utt, seg, frameDim=self.frameDim)
                           for utt, seg in zip(self.fCorpus, self.segmentations)]
        # the reference zips embeddings with captions: the shorter list decides
        n_used = min(len(self.embeddings), len(self.tCorpus))
        self.acoustic_model = self.acoustic_model(
            self.numMixtures, self.frameDim, fCorpus=self.embeddings[:n_used], tCorpus=self.tCorpus[:n_used],
            initProbFile=initProbFile, transProbFile=transProbFile, obsModelFile=obsModelFile,
            initMethod=initMethod, maxLen=maxLen, fixedVariance=fixedVariance)
        print("Finish initialization of acoustic model")

    def embed(self, y, frameDim=None, technique="resample"):
        """One segment (frames x dims) -> (embedDim,) vector (reference :114-143)."""
        feat = y[:, :frameDim].T if frameDim else y.T
        n_out = int(self.embedDim / (frameDim if frameDim else self.featDim))
        if technique == "resample":
            return _resample_rows(feat.astype("float32")[None], n_out)[0].flatten(ORDER)
        if technique == "interpolate":
            grid = np.linspace(0, feat.shape[1] - 1, n_out)
            return interpolate.interp1d(np.arange(feat.shape[1]), feat, kind="linear")(grid).flatten(ORDER)
        if technique == "rasanen":
            whole = (feat.shape[1] // n_out) * n_out
            return feat[:, :whole].reshape((feat.shape[0], n_out, -1)).mean(axis=-1).flatten(ORDER)
        raise UnboundLocalError("local variable 'y_new' referenced before assignment")   # as the reference

    def getSentEmbeds(self, x, segmentation, frameDim=12):
        """All segments of one utterance -> (segments, embedDim).  Segments of equal length go through one
        batched FFT resample; the result equals ``embed`` applied segment by segment."""
        seg = np.asarray(segmentation, dtype=np.int64)
        n_seg = len(seg) - 1
        if n_seg <= 0:
            return np.array([])
        if seg[-1] > len(x) or seg[0] < 0 or np.any(np.diff(seg) <= 0):
            # irregular landmarks (past the utterance end, empty segments): plain slicing semantics
            return np.array([self.embed(x[a:b], frameDim=frameDim) for a, b in zip(seg[:-1], seg[1:])])
        n_out = int(self.embedDim / frameDim)
        lengths = np.diff(seg)
        rows = [None] * n_seg
        for length in np.unique(lengths):
            members = np.flatnonzero(lengths == length)
            take = seg[members, None] + np.arange(int(length))[None, :]              # (m, length) frame ids
            block = np.transpose(x[take][:, :, :frameDim], (0, 2, 1)).astype("float32")   # (m, frameDim, length)
            out = _resample_rows(block, n_out).reshape(len(members), -1)
            for m, row in zip(members, out):
                rows[m] = row
        return np.array(rows)

    def getSentDurations(self, segmentation):
        return np.diff(np.asarray(segmentation)).tolist()

    # ------------------------------------------------------------------ EM / decoding through the acoustic model
    def trainUsingEM(self, numIterations=30, numAMSteps=1, modelPrefix='', writeModel=False):
        am = self.acoustic_model
        if writeModel:
            am.printModel('initial_model.txt')
        for epoch in range(numIterations):
            print("Start training iteration " + str(epoch))
            t0 = time.time()
            am.trainUsingEM(numIterations=numAMSteps)
            print("Acoustic model training takes %0.5f s to finish" % (time.time() - t0))
            if writeModel:
                am.printModel(modelPrefix + "model_iter=" + str(epoch))
        if writeModel:
            am.printModel(modelPrefix + 'model_final')

    def assign(self, i):
        return self.acoustic_model.align(self.embeddings[i], self.tCorpus[i])

    def align(self, i):
        """Per-segment states / scores repeated over each segment's frames.  The acoustic model's scores
        start at the second segment, so -- as in the reference's zip -- align_probs ends one segment
        before the alignment does."""
        dur = np.asarray(self.getSentDurations(self.segmentations[i]), dtype=np.int64)
        states, scores = self.assign(i)
        k = min(len(states), len(dur))
        alignment = np.repeat(np.asarray(states[:k]), dur[:k]).tolist()
        m = min(len(scores), len(dur))
        align_probs = [scores[s] for s in np.repeat(np.arange(m), dur[:m])]
        return alignment, align_probs

    def printAlignment(self, filePrefix, isPhoneme=True):
        n_used = min(len(self.embeddings), len(self.tCorpus))
        records = []
        with open(filePrefix + '.txt', 'w') as txt:
            for i in range(n_used):
                alignment, probs = self.align(i)
                records.append({'index': self.data_ids[i], 'image_concepts': self.tCorpus[i],
                                'alignment': alignment, 'align_probs': probs,
                                'is_phoneme': False, 'is_audio': True})
                txt.write('%s\n%s\n' % (self.tCorpus[i], self.fCorpus[i]))
                txt.write(''.join('%d ' % a for a in alignment) + '\n\n')
        with open(filePrefix + '.json', 'w') as f:
            json.dump(records, f, indent=4, sort_keys=True)
