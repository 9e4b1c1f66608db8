"""BoW quantisation (Database::FrameToBow, src/bow/database.cc:57-89) on the device against the numpy restatement
(oracle/bow.py).  The tree descent compares float squared distances; the reference's own float summation order is unspecified (Eigen), so
word ids are asserted where the oracle's best / second-best margin exceeds float noise and the agreement is reported (it is 100 % here)."""
import os
import struct

import numpy as np
import pytest

import _parity as P

pytestmark = pytest.mark.gpu


def _check(ctx, voc, feats, tag):
    from oracle import bow
    words, bv = ctx.bow_transform(feats)
    w_o, bv_o, margins = bow.transform(voc, feats)
    safe = margins > 1e-5
    P.exact("%s: word ids where the tree decision is not a float near-tie" % tag, np.array_equal(words[safe], w_o[safe]) and safe.mean() > 0.99)
    P.report("%s: word id agreement over all keypoints" % tag, float((words == w_o).mean()), "fraction")
    if np.array_equal(words, w_o):
        P.exact("%s: BowVector word ids (std::map order)" % tag, [a for a, _ in bv] == [a for a, _ in bv_o])
        P.check("%s: BowVector values (idf sums, L1-normalised, double)" % tag, max(abs(x - y) for (_, x), (_, y) in zip(bv, bv_o)), 1e-15)
        assert abs(sum(v for _, v in bv) - 1.0) < 1e-12


def test_bow_words_and_vector_match_oracle():
    from airslam_b200 import capi
    from oracle import bow, synth
    voc = bow.load_vocabulary()
    ctx = capi.Context(max_batch=1, enable_plnet=0, enable_lightglue=0)
    try:
        ctx.bow_load()                                                    # weights/point_voc_L4.afw
        l, r, _ = synth.stereo_pair(752, 480, 601)
        det = ctx.detect_batch(capi.NET_SUPERPOINT, np.stack([l, r]))
        _check(ctx, voc, det[0][0], "BoW on detector features (left)")
        _check(ctx, voc, det[1][0], "BoW on detector features (right)")
        _check(ctx, voc, synth.keypoint_set(400, 752, 480, 5), "BoW on Gaussian descriptors")
        w0, b0 = ctx.bow_transform(np.zeros((259, 0), np.float32))
        assert len(w0) == 0 and b0 == []
    finally:
        ctx.close()


def test_bow_loads_the_reference_boost_archive_layout(tmp_path):
    """airfe_bow_load also reads the reference's own file format (Boost binary archive, include/bow/database.h:35-55).  The real
    voc/point_voc_L4.bin does not travel to the GPU box, so a small tree is written here in exactly that layout."""
    from airslam_b200 import capi
    from oracle import bow
    rs = np.random.RandomState(7)
    k, L = 4, 2
    n = 1 + k + k * k
    children = np.full((n, k), -1, np.int32)
    children[0] = np.arange(1, 1 + k)
    for c in range(k):
        children[1 + c] = 1 + k + c * k + np.arange(k)
    desc = rs.normal(0, 1, (n, 256)).astype(np.float32)
    desc /= np.linalg.norm(desc, axis=1, keepdims=True)
    word_id = np.zeros(n, np.int32)
    word_id[1 + k:] = np.arange(k * k)
    weight = np.zeros(n)
    weight[1 + k:] = rs.uniform(0.5, 3.0, k * k)
    weight[1 + k + 3] = 0.0                                               # a stopped word
    b = bytearray()
    b += struct.pack("<Q", 22) + b"serialization::archive" + struct.pack("<H", 17) + bytes([4, 8, 4, 8]) + struct.pack("<I", 1) + bytes(5)
    b += struct.pack("<4i", k, L, 0, 0) + bytes(5) + struct.pack("<QI", n, 0)
    for i in range(n):
        if i == 0:
            b += bytes([1, 0, 0, 0, 0])
        ch = children[i][children[i] >= 0]
        b += struct.pack("<IId", i, i, weight[i]) + struct.pack("<Q", len(ch)) + ch.astype("<u4").tobytes()
        b += struct.pack("<I", 0) + desc[i].tobytes() + struct.pack("<I", int(word_id[i]))
    path = tmp_path / "tiny_voc.bin"
    path.write_bytes(bytes(b) + bytes(64))
    voc = dict(k=k, L=L, children=children, desc=desc, word_id=word_id, weight=weight)
    feats = np.zeros((259, 64), np.float32)
    d = rs.normal(0, 1, (256, 64))
    feats[3:] = d / np.linalg.norm(d, axis=0, keepdims=True)
    ctx = capi.Context(max_batch=1, enable_superpoint=0, enable_plnet=0, enable_lightglue=0)
    try:
        ctx.bow_load(str(path))
        words, bv = ctx.bow_transform(feats)
        w_o, bv_o, margins = bow.transform(voc, feats)
        P.exact("BoW: Boost-archive loader + transform on a tiny tree", np.array_equal(words, w_o) and [a for a, _ in bv] == [a for a, _ in bv_o] and (w_o == 0xFFFFFFFF).any())
        with pytest.raises(capi.AirfeError):
            (tmp_path / "bad.bin").write_bytes(b"not an archive at all............................")
            ctx.bow_load(str(tmp_path / "bad.bin"))
    finally:
        ctx.close()
