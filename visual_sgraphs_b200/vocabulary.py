"""Python mirror of the part of DBoW2::TemplatedVocabulary that Frame::ComputeBoW / KeyFrame::ComputeBoW use
(reference orb_slam3/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1265, Frame.cc:882-889): the flattened tree on
the device + transform(features, BowVector, FeatureVector, levelsup).  The tree walk runs on the GPU (vsg_bow_transform);
the BowVector / FeatureVector bookkeeping is replayed here in feature order with double arithmetic, as the reference does.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


class Vocabulary:
    """nodes: child lists (CSR: child_ptr, child_idx), node descriptors (nnodes, 32), per-node word_id (-1 for inner
    nodes) and weight; levels = m_L.  TF-IDF weighting with L1 scoring (the ORB vocabulary's settings)."""

    def __init__(self, matcher, child_ptr, child_idx, node_desc, word_id, weight, levels):
        self._L = _lib.load()
        self._matcher = matcher
        self.child_ptr = np.ascontiguousarray(child_ptr, np.int32)
        self.child_idx = np.ascontiguousarray(child_idx, np.int32)
        self.node_desc = np.ascontiguousarray(node_desc, np.uint8).reshape(-1, 32)
        self.word_id = np.ascontiguousarray(word_id, np.int64)
        self.weight = np.ascontiguousarray(weight, np.float64)
        self.levels = int(levels)
        self._h = C.c_void_p()
        check(self._L.vsg_vocabulary_create(matcher._h, len(self.node_desc), ptr(self.child_ptr), ptr(self.child_idx),
                                            ptr(self.node_desc), self.levels, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.vsg_vocabulary_destroy(self._h)
            self._h = None

    __del__ = close

    def walk(self, descriptors, levelsup=4):
        """Per descriptor: (leaf node, FeatureVector node) — the GPU part."""
        d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        leaf = np.zeros(len(d), np.int32)
        nid = np.zeros(len(d), np.int32)
        check(self._L.vsg_bow_transform(self._matcher._h, self._h, ptr(d), len(d), int(levelsup), ptr(leaf), ptr(nid)))
        return leaf, nid

    def transform(self, descriptors, levelsup=4):
        """transform(features, v, fv, levelsup) (TemplatedVocabulary.h:1139-1205), TF_IDF + L1.
        Returns (BowVector as {word: value}, FeatureVector as (nodes, ptr, idx) with sorted node ids)."""
        leaf, nid = self.walk(descriptors, levelsup)
        return bow_bookkeeping(self.word_id[leaf], self.weight[leaf], nid)


def bow_bookkeeping(word_ids, weights, node_ids):
    """BowVector::addWeight / FeatureVector::addFeature in feature order, then BowVector::normalize(L1)."""
    v, fv = {}, {}
    for i, (w_id, w, nd) in enumerate(zip(word_ids.tolist(), weights.tolist(), node_ids.tolist())):
        if w > 0:                                      # not stopped (:1174)
            v[w_id] = v.get(w_id, 0.0) + w             # BowVector::addWeight
            fv.setdefault(nd, []).append(i)            # FeatureVector::addFeature
    norm = 0.0
    for w_id in sorted(v):                             # std::map order (BowVector.cpp normalize, L1)
        norm += abs(v[w_id])
    if norm > 0.0:
        for w_id in v:
            v[w_id] /= norm
    nodes = sorted(fv)
    ptr_, idx = [0], []
    for nd in nodes:
        idx.extend(fv[nd])
        ptr_.append(len(idx))
    return v, (np.array(nodes, np.int32), np.array(ptr_, np.int32), np.array(idx, np.int32))
