"""Ragged / degenerate frames inside one library call (the reference sees them on real sequences: lens cap, saturation, dark corridors --
`FeatureDetector::Detect` is called on whatever the camera delivers, src/map_builder.cc:85):
  * a blank (all-zero) pair, a saturated (all-255) pair, a (textured, blank) pair and a low-contrast pair next to ordinary pairs in ONE call;
  * every pair of that ragged call is BIT-IDENTICAL to the same pair run alone (packed matcher rows: a slot with few or no keypoints must
    not move its neighbours' results), the call is deterministic, and frames are independent (the same left image gives the same features
    whatever its right image is);
  * structural validity: counts within capacity, unit-norm descriptors, match indices inside the two keypoint sets, one-to-one matches,
    scores in (threshold, 1]; a side without keypoints gives zero matches (src/point_matcher.cc:43-48 early return)."""
import numpy as np
import pytest

import _parity as P

pytestmark = pytest.mark.gpu


def _same(a, b):
    return (np.array_equal(a["feat_l"], b["feat_l"]) and np.array_equal(a["feat_r"], b["feat_r"]) and np.array_equal(a["lines_l"], b["lines_l"]) and
            np.array_equal(a["lines_r"], b["lines_r"]) and np.array_equal(a["junc"], b["junc"]) and np.array_equal(a["matches"][0], b["matches"][0]) and
            np.array_equal(a["matches"][1], b["matches"][1]))


@pytest.fixture(scope="module")
def ragged():
    from airslam_b200 import capi
    from oracle import synth
    W, H = 752, 480
    a_l, a_r, _ = synth.stereo_pair(W, H, 0xA1750002)
    b_l, b_r, _ = synth.stereo_pair(W, H, 0xA1750002 + 99)
    blank = np.zeros((H, W), np.uint8)
    white = np.full((H, W), 255, np.uint8)
    low = (128 + (a_l.astype(np.int32) - 128) // 32).astype(np.uint8)          # 8 grey levels around mid-grey: few keypoints survive the threshold
    left = np.stack([a_l, blank, a_l, white, low, b_l])
    right = np.stack([a_r, blank, blank, white, low, b_r])
    ctx = capi.Context(max_batch=len(left), enable_superpoint=0)
    out = ctx.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, left, right, lines=True, junctions=True)
    yield ctx, left, right, out
    ctx.close()


def test_ragged_call_is_batch_invariant_and_deterministic(ragged):
    from airslam_b200 import capi
    ctx, left, right, out = ragged
    counts = [(o["feat_l"].shape[1], o["feat_r"].shape[1], len(o["matches"][0])) for o in out]
    P.report("ragged call: (left keypoints, right keypoints, matches) per pair", float(len(out)), "pairs", str(counts))
    for k in range(len(left)):
        alone = ctx.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, left[k:k + 1], right[k:k + 1], lines=True, junctions=True)[0]
        P.exact("ragged call: pair %d == alone (all outputs, bitwise)" % k, _same(out[k], alone))
    again = ctx.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, left, right, lines=True, junctions=True)
    P.exact("ragged call: deterministic", all(_same(a, b) for a, b in zip(out, again)))
    # frame independence: pairs 0 and 2 share the left image
    P.exact("ragged call: same left image, different right image -> same left features / lines", np.array_equal(out[0]["feat_l"], out[2]["feat_l"]) and np.array_equal(out[0]["lines_l"], out[2]["lines_l"]))
    P.exact("ragged call: the blank right image of pair 2 == the blank frames of pair 1", np.array_equal(out[2]["feat_r"], out[1]["feat_r"]) and np.array_equal(out[1]["feat_l"], out[1]["feat_r"]))


def test_ragged_call_outputs_are_structurally_valid(ragged):
    _, left, _, out = ragged
    for k, o in enumerate(out):
        for side in ("feat_l", "feat_r"):
            f = o[side]
            assert f.shape[0] == 259 and f.shape[1] <= 400, (k, side, f.shape)
            if f.shape[1]:
                assert np.isfinite(f).all(), (k, side)
                assert (f[1] >= 0).all() and (f[1] <= 752).all() and (f[2] >= 0).all() and (f[2] <= 480).all(), (k, side)
                nrm = np.linalg.norm(f[3:], axis=0)
                ok = (np.abs(nrm - 1.0) < 1e-4) | (nrm == 0.0)                       # extract_descriptors: unit norm, or the zero column of a zero sample
                assert ok.all(), (k, side, float(np.abs(nrm - 1).max()))
        idx, sc = o["matches"]
        nl, nr = o["feat_l"].shape[1], o["feat_r"].shape[1]
        if nl == 0 or nr == 0:
            P.exact("ragged call: pair %d has an empty side -> zero matches" % k, len(idx) == 0)
            continue
        assert idx.shape[0] == sc.shape[0]
        if len(idx):
            assert (idx[:, 0] >= 0).all() and (idx[:, 0] < nl).all() and (idx[:, 1] >= 0).all() and (idx[:, 1] < nr).all(), k
            assert len(set(idx[:, 0].tolist())) == len(idx) and len(set(idx[:, 1].tolist())) == len(idx), k      # mutual nearest neighbours: one-to-one
            assert np.isfinite(sc).all() and (sc > 0.1 - 1e-6).all() and (sc <= 1.0 + 1e-6).all(), (k, float(sc.min()), float(sc.max()))
        assert o["lines_l"].ndim == 2 and o["lines_l"].shape[1] == 4 and np.isfinite(o["lines_l"]).all(), k
    # the two ordinary pairs still match well
    for k in (0, 5):
        assert len(out[k]["matches"][0]) > 100, (k, len(out[k]["matches"][0]))
