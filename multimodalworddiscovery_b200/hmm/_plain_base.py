"""Shared host-side logic of the plain-state HMM word discoverers (hmm/ classes).

Mirrors hmm/hmm_word_discoverer.py (HMMWordDiscoverer) and hmm/audio_hmm_word_discoverer.py
(AudioHMMWordDiscoverer): same constructor / method names / public attributes / files.  The
dict-of-dict ``obs`` is kept dense on the device; ``self.obs`` materialises the reference's dict view.
"""
import json

import numpy as np

from ..engine_hmm import PackedSentences, PlainHMMEngine

NULL = "NULL"


class PlainHMMBase(object):
    LOG = False

    # ------------------------------------------------------------------ corpus
    def _read_blocks(self, fileName, add_null):
        """3-line blocks 'concepts / phones / blank' (hmm_word_discoverer.py:34-55)."""
        f = open(fileName)
        i = 0
        for s in f:
            if i == 0:
                tTokenized = s.split()
                if add_null:
                    tTokenized.insert(0, NULL)                 # audio_hmm_word_discoverer.py:52
                self.tCorpus.append(tTokenized)
            elif i == 1:
                self.fCorpus.append(s.split())
            else:
                i = -1
            i += 1
        f.close()

    def computeTranslationLengthProbabilities(self, smoothing=None):
        """:206-237 incl. the per-sentence reset of lenProb[len(ts)]."""
        for ts, fs in zip(self.tCorpus, self.fCorpus):
            self.lenProb[len(ts)] = {}
            if len(fs) not in self.lenProb[len(ts)].keys():
                self.lenProb[len(ts)][len(fs)] = 1
            else:
                self.lenProb[len(ts)][len(fs)] += 1
        if smoothing == 'laplace':
            tLenMax = max(list(self.lenProb.keys()))
            fLenMax = max([max(list(f.keys())) for f in list(self.lenProb.values())])
            for tLen in range(tLenMax):
                for fLen in range(fLenMax):
                    if tLen not in self.lenProb:
                        self.lenProb[tLen] = {}
                        self.lenProb[tLen][fLen] = 1.
                    elif fLen not in self.lenProb[tLen]:
                        self.lenProb[tLen][fLen] = 1.
                    else:
                        self.lenProb[tLen][fLen] += 1.
        for tl in self.lenProb.keys():
            totCount = sum(self.lenProb[tl].values())
            for fl in self.lenProb[tl].keys():
                self.lenProb[tl][fl] = self.lenProb[tl][fl] / totCount

    # ------------------------------------------------------------------ vocab / dense obs
    def _vocab(self):
        token = (id(self.tCorpus), len(self.tCorpus), id(self.fCorpus), len(self.fCorpus))
        if getattr(self, '_vocab_token', None) != token:
            self._tv = {w: i for i, w in enumerate(sorted({w for e in self.tCorpus for w in e}))}
            self._fv = {w: i for i, w in enumerate(sorted({w for s in self.fCorpus for w in s}))}
            self._tw = sorted(self._tv)
            self._fw = sorted(self._fv)
            self._tgt_ids = [np.array([self._tv[w] for w in e], dtype=np.int32) for e in self.tCorpus]
            self._src_ids = [np.array([self._fv[w] for w in s], dtype=np.int32) for s in self.fCorpus]
            self._vocab_token = token
            self._eng = None
        return self._tv, self._fv

    def _dist(self):
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                return dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass
        return 0, 1

    def _engine(self):
        self._vocab()
        if getattr(self, '_eng', None) is None:
            rank, world = self._dist()
            pk = PackedSentences(self._tgt_ids, self._src_ids, len(self._fv), rank=rank, world=world)
            self._eng = PlainHMMEngine(pk, len(self._tv), len(self._fv), self.LOG,
                                       device=getattr(self, '_device', None))
        return self._eng

    def _dict_to_dense(self, obs):
        tv, fv = self._vocab()
        d = np.full((len(tv), len(fv)), np.nan)
        for tw, row in obs.items():
            if tw not in tv:
                continue
            for fw, v in row.items():
                if fw in fv:
                    d[tv[tw], fv[fw]] = v
        return d

    def _dense_to_dict(self, dense):
        out = {}
        rows, cols = np.nonzero(~np.isnan(dense))
        for r, c in zip(rows.tolist(), cols.tolist()):
            out.setdefault(self._tw[r], {})[self._fw[c]] = float(dense[r, c])
        return out

    @property
    def obs(self):
        if self._obs_dict is None and self._obs_dense is not None:
            self._obs_dict = self._dense_to_dict(self._obs_dense)
        return self._obs_dict

    @obs.setter
    def obs(self, v):
        self._obs_dict = v
        self._obs_dense = None

    def _push(self):
        eng = self._engine()
        if self._obs_dense is None:
            self._obs_dense = self._dict_to_dense(self._obs_dict or {})
        eng.set_params(self.init, self.trans, self._obs_dense)
        return eng

    def _pull(self, eng):
        init, trans, dense = eng.get_params()
        for m in init:
            self.init[m] = init[m]
            self.trans[m] = trans[m]
        self._obs_dense = dense
        self._obs_dict = None

    def _load_param_files(self):
        if self.initProbFile:
            with open(self.initProbFile) as f:
                for line in f:
                    m, s, prob = line.split()
                    self.init[int(m)][int(s)] = float(prob)
        if self.transProbFile:
            with open(self.transProbFile) as f:
                for line in f:
                    m, cur_s, next_s, prob = line.split()
                    self.trans[int(m)][int(cur_s)][int(next_s)] = float(prob)
        if self.obsProbFile:
            obs = {}
            with open(self.obsProbFile) as f:
                for line in f:
                    tw, fw, prob = line.strip().split()
                    obs.setdefault(tw, {})[fw] = float(prob)
            return obs
        return None

    # ------------------------------------------------------------------ single-pair API
    def _pair_ids(self, eSen, fSen):
        tv, fv = self._vocab()
        return (np.array([tv[w] for w in eSen], dtype=np.int32), np.array([fv[w] for w in fSen], dtype=np.int32))

    def _mini(self, eSen, fSen):
        e, f = self._pair_ids(eSen, fSen)
        self._push()
        pk = PackedSentences([e], [f], len(self._fv))
        pk.lens = self._eng.pk.lens
        eng = PlainHMMEngine(pk, len(self._tv), len(self._fv), self.LOG, device=getattr(self, '_device', None))
        eng.init_t.copy_(self._eng.init_t)
        eng.trans_t.copy_(self._eng.trans_t)
        eng.obs.copy_(self._eng.obs)
        return eng, len(e), len(f)

    def forward(self, eSen, fSen):
        eng, n, T = self._mini(eSen, fSen)
        al, _ = eng.dense_sweeps()
        return al.cpu().numpy()[:n * T].reshape(T, n)

    def backward(self, eSen, fSen):
        eng, n, T = self._mini(eSen, fSen)
        _, be = eng.dense_sweeps()
        return be.cpu().numpy()[:n * T].reshape(T, n)

    def align(self, fSen, eSen, unkProb=10e-12):
        eng, n, T = self._mini(eSen, fSen)
        ali, ap = eng.align(unkProb)
        return [int(a) for a in ali.cpu().numpy()], ap.cpu().numpy().reshape(T - 1, n).tolist()

    def computeAvgLogLikelihood(self):
        eng = self._push()
        return float(eng.loglik_sum()) / len(self.tCorpus)

    # ------------------------------------------------------------------ I/O
    def printModel(self, fileName):
        """:331-352"""
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
        obsFile = open(fileName + '_observationprobs.txt', 'w')
        obs = self.obs or {}
        for tw in sorted(obs):
            for fw in sorted(obs[tw]):
                obsFile.write('%s\t%s\t%f\n' % (tw, fw, obs[tw][fw]))
        obsFile.close()

    def _align_all(self):
        eng = self._push()
        ali, ap = eng.align()
        ali, ap = ali.cpu().numpy(), ap.cpu().numpy()
        pk = eng.pk
        local = [(int(ex), ali[pk.src_off[s]:pk.src_off[s + 1]], ap[pk.ap_off[s]:pk.ap_off[s + 1]])
                 for s, ex in enumerate(pk.order)]
        rank, world = self._dist()
        if world > 1:
            import torch.distributed as dist
            allp = [None] * world
            dist.all_gather_object(allp, local)
            local = [x for part in allp for x in part]
        out = [None] * len(self.tCorpus)
        for ex, a, p in local:
            out[ex] = (a, p)
        return out

    def _print_alignment(self, filePrefix, extra):
        res = self._align_all()
        rank, _ = self._dist()
        if rank != 0:
            return
        f = open(filePrefix + '.txt', 'w')
        aligns = []
        for i, (fSen, tSen) in enumerate(zip(self.fCorpus, self.tCorpus)):
            a, p = res[i]
            info = {'index': i, 'image_concepts': tSen, 'alignment': [int(x) for x in a],
                    'align_probs': np.asarray(p).reshape(-1, len(tSen)).tolist()}
            info.update(extra(fSen))
            aligns.append(info)
            f.write('%s\n%s\n' % (tSen, fSen))
            for x in a:
                f.write('%d ' % x)
            f.write('\n\n')
        f.close()
        with open(filePrefix + '.json', 'w') as f:
            json.dump(aligns, f, indent=4, sort_keys=True)
