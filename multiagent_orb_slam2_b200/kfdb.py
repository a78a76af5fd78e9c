"""Keyframe database for place recognition: KeyFrameDatabase::add / erase / DetectLoopCandidates
(/root/reference/src/KeyFrameDatabase.cc:40-66, 76-197) with the word-sharing counts and the DBoW2 L1 scores of the query
against every keyframe computed on the device in one launch (orbdb_query); the candidate bookkeeping that walks the
covisibility graph is replayed on the host from those arrays."""
import ctypes as C

import numpy as np

from . import _lib


class KeyFrameDatabase:
    def __init__(self, n_words, device=0):
        self._L = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._L.orbdb_create(device, int(n_words), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbdb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self._L.orbdb_size(self._h)

    def add(self, word_ids, weights):
        """KeyFrameDatabase::add (40-46): the keyframe's BowVector (ascending word ids, weights); returns its slot."""
        ids = np.ascontiguousarray(word_ids, np.int32); w = np.ascontiguousarray(weights, np.float64)
        slot = C.c_int()
        _lib.check(self._L.orbdb_add(self._h, ids.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), len(ids), C.byref(slot)))
        return slot.value

    def erase(self, slot):
        _lib.check(self._L.orbdb_erase(self._h, int(slot)))

    def score_all(self, word_ids, weights):
        """(common words, first shared word, L1 score) of the query BowVector against every slot."""
        ids = np.ascontiguousarray(word_ids, np.int32); w = np.ascontiguousarray(weights, np.float64)
        n = len(self)
        common = np.zeros(n, np.int32); first = np.full(n, -1, np.int32); score = np.zeros(n, np.float64)
        _lib.check(self._L.orbdb_query(self._h, ids.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), len(ids),
                                       common.ctypes.data_as(C.c_void_p), first.ctypes.data_as(C.c_void_p), score.ctypes.data_as(C.c_void_p), n))
        return common, first, score

    def DetectLoopCandidates(self, word_ids, weights, min_score, connected=(), neighbours=lambda slot: ()):
        """KeyFrameDatabase::DetectLoopCandidates (76-197). connected: slots of pKF->GetConnectedKeyFrames();
        neighbours(slot): slots of GetBestCovisibilityKeyFrames(10). Returns candidate slots in the reference's order."""
        f32 = np.float32
        common, first, score = self.score_all(word_ids, weights)
        connected = set(int(c) for c in connected)
        share = [s for s in np.flatnonzero(common > 0) if int(s) not in connected]
        if not share:
            return []
        # the reference meets the keyframes word by word (ascending) and, inside a word's list, in insertion order
        share.sort(key=lambda s: (first[s], s))
        in_list = set(int(s) for s in share)
        min_common = int(f32(max(int(common[s]) for s in share)) * f32(0.8))
        si = score.astype(f32)
        matches = [(si[s], int(s)) for s in share if common[s] > min_common and si[s] >= f32(min_score)]
        if not matches:
            return []
        acc, best_acc = [], f32(min_score)
        for s0, slot in matches:
            best_score, acc_score, best_kf = s0, s0, slot
            for nb in neighbours(slot):
                nb = int(nb)
                if nb in in_list and common[nb] > min_common:
                    acc_score = f32(acc_score + si[nb])
                    if si[nb] > best_score:
                        best_kf, best_score = nb, si[nb]
            acc.append((acc_score, best_kf))
            if acc_score > best_acc:
                best_acc = acc_score
        retain = f32(f32(0.75) * best_acc)
        out, seen = [], set()
        for a, kf in acc:
            if a > retain and kf not in seen:
                out.append(kf); seen.add(kf)
        return out

    # ---- the two other place-recognition queries of the reference, same device scoring ------------------------------------
    def _sharing(self, word_ids, weights, skip=()):
        common, first, score = self.score_all(word_ids, weights)
        skip = set(int(c) for c in skip)
        share = [int(s) for s in np.flatnonzero(common > 0) if int(s) not in skip]
        share.sort(key=lambda s: (first[s], s))   # ascending word, then insertion order of the word's list
        return share, common, score.astype(np.float32)

    @staticmethod
    def _accumulate(matches, in_list, gate, stale, neighbours, best_acc):
        f32 = np.float32
        acc = []
        for s0, slot in matches:
            best_score, acc_score, best_kf = s0, s0, slot
            for nb in neighbours(slot):
                nb = int(nb)
                if nb in in_list and gate(nb):
                    acc_score = f32(acc_score + f32(stale[nb]))
                    if f32(stale[nb]) > best_score:
                        best_kf, best_score = nb, f32(stale[nb])
            acc.append((acc_score, best_kf))
            if acc_score > best_acc:
                best_acc = acc_score
        retain = f32(f32(0.75) * best_acc)
        out, seen = [], set()
        for a, kf in acc:
            if a > retain and kf not in seen:
                out.append(kf); seen.add(kf)
        return out

    def DetectCovisibilityCandidates(self, word_ids, weights, min_score, ignore=(), neighbours=lambda slot: (), covis_score=None):
        """KeyFrameDatabase::DetectCovisibilityCandidates (199-308), the fork's query behind MapFusion::CovisibilityDiscovery.
        The reference never stores the similarity in the keyframe there (mCovisScore is only read, 270-276): the covisibility
        accumulation adds what the keyframes held before - covis_score[slot], zeros by default."""
        f32 = np.float32
        share, common, si = self._sharing(word_ids, weights, ignore)
        if not share:
            return []
        stale = np.zeros(len(self), f32) if covis_score is None else np.asarray(covis_score, f32)
        min_common = int(f32(max(int(common[s]) for s in share)) * f32(0.8))
        matches = [(si[s], s) for s in share if common[s] > min_common and si[s] >= f32(min_score)]
        if not matches:
            return []
        return self._accumulate(matches, set(share), lambda nb: common[nb] > min_common, stale, neighbours, f32(min_score))

    def DetectRelocalizationCandidates(self, word_ids, weights, neighbours=lambda slot: (), reloc_score=None):
        """KeyFrameDatabase::DetectRelocalizationCandidates (310-420). reloc_score: per-slot mRelocScore, updated in place
        (a neighbour below the common-word floor contributes the score of an earlier query, 378-387); kept on the object
        when not given."""
        f32 = np.float32
        if reloc_score is None:
            if getattr(self, "_reloc", None) is None or len(self._reloc) < len(self):
                old = getattr(self, "_reloc", None)
                self._reloc = np.zeros(len(self), f32)
                if old is not None:
                    self._reloc[:len(old)] = old
            reloc_score = self._reloc
        share, common, si = self._sharing(word_ids, weights)
        if not share:
            return []
        min_common = int(f32(max(int(common[s]) for s in share)) * f32(0.8))
        matches = []
        for s in share:
            if common[s] > min_common:
                reloc_score[s] = si[s]
                matches.append((si[s], s))
        if not matches:
            return []
        return self._accumulate(matches, set(share), lambda nb: True, reloc_score, neighbours, f32(0))
