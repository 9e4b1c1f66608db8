"""The bench configuration itself (32 stereo pairs per call, PLNet + LightGlue, 752x480) under parity:
  * batch invariance: pair k inside a 32-pair call is BIT-IDENTICAL to the same pair run alone -- every output (features, lines,
    junctions, match indices and scores).  This holds because a tile's K accumulation order never depends on the CTA that gets it
    (tc_ffn.cuh used to rotate the weight-tile order per CTA; removed) and every reduction is per row / per image;
  * determinism: the same 32-pair call twice gives identical bytes;
  * oracle: LightGlue on the GPU's own features of pairs 0 and 31 gives the oracle's match indices;
  * domain properties at full size: median disparity of the matches == the planted disparity of every pair."""
import numpy as np
import pytest

import _parity as P

pytestmark = pytest.mark.gpu

NP = 32


@pytest.fixture(scope="module")
def run32():
    from airslam_b200 import capi
    from oracle import synth
    ims = [synth.stereo_pair(752, 480, 0xA1750002 + 7 * k) for k in range(NP)]
    left = np.stack([a for a, _, _ in ims])
    right = np.stack([b for _, b, _ in ims])
    ctx = capi.Context(max_batch=NP, enable_superpoint=0)
    out = ctx.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, left, right, lines=True, junctions=True)
    yield ctx, left, right, [d for _, _, d in ims], out
    ctx.close()


def _same(a, b):
    return (np.array_equal(a["feat_l"], b["feat_l"]) and np.array_equal(a["feat_r"], b["feat_r"]) and np.array_equal(a["lines_l"], b["lines_l"]) and
            np.array_equal(a["lines_r"], b["lines_r"]) and np.array_equal(a["junc"], b["junc"]) and np.array_equal(a["matches"][0], b["matches"][0]) and
            np.array_equal(a["matches"][1], b["matches"][1]))


def test_pair_inside_batch32_equals_pair_alone(run32):
    from airslam_b200 import capi
    ctx, left, right, _, out = run32
    for k in (0, 13, 31):
        alone = ctx.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, left[k:k + 1], right[k:k + 1], lines=True, junctions=True)[0]
        P.exact("batch invariance: pair %d of 32 == alone (all outputs, bitwise)" % k, _same(out[k], alone))
    # a sub-batch starting elsewhere: slots move, results must not
    sub = ctx.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, left[5:13], right[5:13], lines=True, junctions=True)
    P.exact("batch invariance: pairs 5..12 as an 8-batch == inside the 32-batch", all(_same(out[5 + i], sub[i]) for i in range(8)))


def test_batch32_is_deterministic(run32):
    from airslam_b200 import capi
    ctx, left, right, _, out = run32
    again = ctx.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, left, right, lines=True, junctions=True)
    P.exact("determinism: 32-pair call twice", all(_same(a, b) for a, b in zip(out, again)))


def test_batch32_matches_equal_oracle_on_own_features(run32):
    from oracle import host, weights
    _, _, _, _, out = run32
    w = weights.load("lightglue")
    try:      # offline analysis aid: the GPU's features / matches of pair 0 travel back with gpurun_out/
        import os
        np.savez(os.path.join(P.ROOT, "gpurun_out", "debug_p32_pair0.npz"), feat_l=out[0]["feat_l"], feat_r=out[0]["feat_r"], idx=out[0]["matches"][0], score=out[0]["matches"][1])
    except Exception:
        pass
    for k in (0, 31):
        # gate: the kernel-matched oracle mode (fp16 rounding of the attention probabilities where tc_attn.cuh does it); report: plain emul
        # (profiles/r02_attention_rounding_drift.txt explains why those two differ by ~1e-2 on ambiguous matches of real-image features)
        m = host.matching_points(out[k]["feat_l"], out[k]["feat_r"], w, 0, 752, 480, emul="fused")
        idx_o = np.array([[a, b] for a, b, _ in m], dtype=np.int32).reshape(-1, 2)
        P.exact("P=32: LightGlue indices of pair %d == kernel-matched oracle on the same features" % k, np.array_equal(out[k]["matches"][0], idx_o))
        sc_o = np.array([1.0 - d for _, _, d in m], dtype=np.float32)
        # Real-image feature sets contain a few AMBIGUOUS matches (probability ~0.5) whose score is chaotic in the fp32 summation order and the
        # fp16 rounding pattern: on pair 0 match #268 reads 0.5716 (fused oracle) / 0.5498 (emul) / 0.5704 (fp32) -- the ORACLE's own three
        # arithmetic modes spread by 2.2e-2 -- while the other 366 matches agree to 4e-5 in every mode, and the GPU value moved between
        # 0.5667, 0.5587 and 0.5532 across round-2 kernel versions that only re-ordered fp32 sums (profiles/r02_attention_rounding_drift.txt).
        # Gates: median within 2e-6 and >= 99 % of the matches within 1e-3 of the kernel-matched mode; EVERY match within 1e-3 of the envelope
        # [min, max] of the oracle's three modes (a kernel bug would leave the envelope; a re-ordered sum cannot).  The raw worst distance to the
        # fused mode is reported.
        err = np.abs(out[k]["matches"][1] - sc_o)
        P.check("P=32: LightGlue scores vs kernel-matched oracle: median", np.median(err), 2e-6, "abs in probability")
        P.check("P=32: LightGlue scores vs kernel-matched oracle: fraction of matches off by > 1e-3", float((err > 1e-3).mean()), 0.01, "fraction")
        P.report("P=32: LightGlue scores vs kernel-matched oracle: worst (ambiguous) match", err.max(), "abs in probability",
                 "reported: chaotic on ambiguous matches, gated through the oracle-mode envelope below")
        modes = [sc_o]
        for em in (True, False):
            mm_ = host.matching_points(out[k]["feat_l"], out[k]["feat_r"], w, 0, 752, 480, emul=em)
            if np.array_equal(np.array([[a, b] for a, b, _ in mm_], dtype=np.int32).reshape(-1, 2), idx_o):
                modes.append(np.array([1.0 - d for _, _, d in mm_], dtype=np.float32))
        lo, hi = np.min(modes, axis=0), np.max(modes, axis=0)
        g = out[k]["matches"][1]
        outside = np.maximum(np.maximum(lo - g, g - hi), 0.0)
        P.check("P=32: LightGlue scores: worst distance to the envelope of the oracle's fp32 / emul / fused modes", outside.max(), 1e-3, "abs in probability",
                "%d oracle modes with identical indices; envelope width max %.3e" % (len(modes), float((hi - lo).max())))
        m2 = host.matching_points(out[k]["feat_l"], out[k]["feat_r"], w, 0, 752, 480, emul=True)
        idx_2 = np.array([[a, b] for a, b, _ in m2], dtype=np.int32).reshape(-1, 2)
        P.exact("P=32: LightGlue indices of pair %d == plain emul oracle" % k, np.array_equal(out[k]["matches"][0], idx_2))
        P.report("P=32: LightGlue score drift vs plain emul oracle (normalised-P rounding)", np.abs(out[k]["matches"][1] - np.array([1.0 - d for _, _, d in m2], dtype=np.float32)).max(),
                 "abs in probability", "reported: see profiles/r02_attention_rounding_drift.txt")


def test_batch32_stereo_geometry(run32):
    _, _, _, disp, out = run32
    for k in range(NP):
        i0, i1 = out[k]["matches"][0][:, 0], out[k]["matches"][0][:, 1]
        assert len(i0) > 100, (k, len(i0))
        med = float(np.median(out[k]["feat_l"][1, i0] - out[k]["feat_r"][1, i1]))
        P.check("P=32: |median match disparity - planted disparity|", abs(med - disp[k]), 1.5, "px")
        assert out[k]["feat_l"].shape[1] == 400 and out[k]["junc"].shape[0] == 259 and out[k]["lines_l"].shape[1] == 4
