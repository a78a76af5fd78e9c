"""Host-side mirror of ORB_SLAM2::ORBVocabulary (= DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>,
/root/reference/include/ORBVocabulary.h) for the one call on the frontend path:
transform(features, BowVector&, FeatureVector&, levelsup) as used by Frame::ComputeBoW (src/Frame.cc:395-402).
The tree descent (all Hamming work) runs on the device (orbv_descend); the BowVector / FeatureVector are
assembled on the host in the reference's order so that the double sums are bit-identical."""
import ctypes as C

import numpy as np

from . import _lib


class ORBVocabulary:
    def __init__(self, voc=None, device=0):
        self._L = _lib.lib()
        self._h = None
        self.device = int(device)
        if voc is not None:
            self.set_nodes(voc["parent"], voc["desc"], voc["weight"], voc["k"], voc["L"])

    def set_nodes(self, parent, desc, weight, k, L):
        parent = np.ascontiguousarray(parent, np.int32)
        desc = np.ascontiguousarray(desc, np.uint8)
        weight = np.ascontiguousarray(weight, np.float64)
        self.close()
        h = C.c_void_p()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(self._L.orbv_create(self.device, p(parent), p(desc), p(weight), len(parent), int(k), int(L), C.byref(h)))
        self._h, self.k, self.Lv = h, int(k), int(L)

    def loadFromTextFile(self, path):
        """TemplatedVocabulary::loadFromTextFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1424)."""
        with open(path) as f:
            k, L, _scoring, _weighting = [int(x) for x in f.readline().split()]
            rows = [l.split() for l in f if l.strip()]
        parent = np.array([-1] + [int(r[0]) for r in rows], np.int32)
        desc = np.zeros((len(rows) + 1, 32), np.uint8)
        desc[1:] = np.array([[int(x) for x in r[2:34]] for r in rows], np.uint8)
        weight = np.array([0.0] + [float(r[34]) for r in rows], np.float64)
        self.set_nodes(parent, desc, weight, k, L)
        return True

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def descend(self, descriptors, levelsup=4):
        d = np.ascontiguousarray(descriptors, np.uint8)
        n = len(d)
        word = np.empty(n, np.int32); node = np.empty(n, np.int32); weight = np.empty(n, np.float64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(self._L.orbv_descend(self._h, p(d), n, int(levelsup), p(word), p(node), p(weight)))
        return word, node, weight

    def transform(self, descriptors, levelsup=4):
        """Returns (bow_ids, bow_values, feature_vector): the BowVector as ascending word ids + L1-normalised
        TF-IDF values, the FeatureVector as a list of (node id, [feature indices]) in ascending node order."""
        word, node, weight = self.descend(descriptors, levelsup)
        bow, fv = {}, {}
        for i in range(len(word)):  # TemplatedVocabulary.h:1150-1163, in feature order
            w = float(weight[i])
            if w > 0:
                wid = int(word[i])
                bow[wid] = bow[wid] + w if wid in bow else w      # BowVector::addWeight
                fv.setdefault(int(node[i]), []).append(i)       # FeatureVector::addFeature
        ids = sorted(bow)
        norm = 0.0
        for wid in ids:                                          # BowVector::normalize(L1), map order
            norm += abs(bow[wid])
        vals = np.array([bow[w] / norm if norm > 0.0 else bow[w] for w in ids], np.float64)
        return np.array(ids, np.int32), vals, sorted(fv.items())
