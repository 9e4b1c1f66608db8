"""AssignPointsToLines + MatchLines (src/line_processor.cc:68-187) with the stereo filter of Frame::AddRightFeatures
(src/frame.cc:141-155) on the device-resident results of a stereo call -- bit-exact against the numpy restatement (integer outputs and
float distances; the arithmetic is double precision on both sides)."""
import numpy as np
import pytest

import _parity as P

pytestmark = pytest.mark.gpu


def test_stereo_line_assoc_matches_oracle():
    from airslam_b200 import capi
    from oracle import host, synth
    ims = [synth.stereo_pair(752, 480, 501 + k) for k in range(3)]
    left = np.stack([a for a, _, _ in ims])
    right = np.stack([b for _, b, _ in ims])
    ctx = capi.Context(max_batch=3, enable_superpoint=0, line_threshold=0.5, line_length_threshold=20.0)      # more, shorter lines: denser association
    try:
        out = ctx.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, left, right, lines=True, junctions=True)
        mn, mx, my = 2.0, 120.0, 3.0
        rn, ri, rd, lm = ctx.stereo_line_assoc(3, mn, mx, my)
        tot_rel = tot_lm = 0
        for p in range(3):
            o = out[p]
            rel_l = host.assign_points_to_lines(o["lines_l"], o["feat_l"])
            rel_r = host.assign_points_to_lines(o["lines_r"], o["feat_r"])
            for side, rel in ((0, rel_l), (1, rel_r)):
                s = 2 * p + side
                got = [{int(ri[s, i, k]): float(rd[s, i, k]) for k in range(rn[s, i])} for i in range(len(rel))]
                P.exact("AssignPointsToLines: point sets and float distances per line (pair %d, %s)" % (p, "LR"[side]),
                        got == rel and [list(g) for g in got] == [sorted(g) for g in got] and rn[s, len(rel):].sum() == 0)
                tot_rel += sum(len(g) for g in rel)
            m = host.filter_stereo_matches(o["feat_l"], o["feat_r"], o["matches"][0], mn, mx, my)
            exp = host.match_lines(rel_l, rel_r, m, o["feat_l"].shape[1], o["feat_r"].shape[1])
            P.exact("MatchLines: line_matches (pair %d)" % p, lm[p, :len(exp)].tolist() == exp and (lm[p, len(exp):] == -1).all())
            tot_lm += sum(1 for v in exp if v >= 0)
        assert tot_rel > 50 and tot_lm > 5, (tot_rel, tot_lm)     # the scene must exercise the code: points on lines, matched stereo lines
    finally:
        ctx.close()
