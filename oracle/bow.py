"""Oracle restatement of the BoW quantisation of keyframe descriptors.  TEST INFRASTRUCTURE.

Database::FrameToBow, src/bow/database.cc:57-89 -> TemplatedVocabulary::transform(feature, word_id, weight)
(3rdparty/DBoW2/include/DBoW2/TemplatedVocabulary.h:1313-1351): descend the k-ary tree, at every level pick the child with the smallest
FSuperpoint::distance = |a - b|^2 (src/bow/FSuperpoint.cc:46-50; strict '<' scan: the first minimum wins), the leaf gives (word id, idf weight);
BowVector::addWeight accumulates per word in feature order and BowVector::normalize(L1) divides by the sum (3rdparty/DBoW2/src/BowVector.cpp:34-84).
The reference evaluates the squared distance in float with Eigen's (vectorised, order-unspecified) reduction; this restatement uses float64, and
reports the margin between the best and the second-best child so that tests can tell a genuine difference from a float near-tie."""
import os

import numpy as np

from . import weights


def load_vocabulary():
    t = weights.load_container(os.path.join(weights.WEIGHT_DIR, "point_voc_L4.afw"), as_float32=False)
    k, L, weighting, scoring = (int(x) for x in t["voc.meta"])
    n = t["voc.children"].shape[0]
    return dict(k=k, L=L, weighting=weighting, scoring=scoring, children=t["voc.children"], desc=t["voc.desc"].reshape(n, 256),
                word_id=t["voc.word_id"], weight=t["voc.weight_f64_bits"].reshape(-1).view(np.float64))


def transform(voc, feat259):
    """feat259 [259, N] -> (word_of_feature uint32 [N] (0xFFFFFFFF = stopped word), bow: sorted list of (word id, value), margins [N])."""
    n = feat259.shape[1]
    words = np.zeros(n, np.uint32)
    margins = np.full(n, np.inf)
    bow = {}
    for i in range(n):
        f = feat259[3:, i].astype(np.float64)
        node = 0
        while voc["children"][node, 0] >= 0:
            ch = voc["children"][node]
            ch = ch[ch >= 0]
            d = ((f[None, :] - voc["desc"][ch].astype(np.float64)) ** 2).sum(1)
            best = int(np.argmin(d))                      # first minimum
            if len(d) > 1:
                margins[i] = min(margins[i], float(np.partition(d, 1)[1] - d[best]))
            node = int(ch[best])
        w = float(voc["weight"][node])
        if w > 0:
            wid = int(voc["word_id"][node])
            bow[wid] = bow.get(wid, 0.0) + w               # BowVector::addWeight, feature order
            words[i] = wid
        else:
            words[i] = 0xFFFFFFFF
    ids = sorted(bow)
    vals = [bow[k] for k in ids]
    norm = 0.0
    for v in vals:                                         # BowVector::normalize(L1): map order
        norm += abs(v)
    if norm > 0:
        vals = [v / norm for v in vals]
    return words, list(zip(ids, vals)), margins
