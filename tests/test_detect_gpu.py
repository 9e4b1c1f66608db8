"""Stage-wise + end-to-end GPU parity of the detector (SuperPoint G1, PLNet G2+G3) against the oracle.

Every dense tolerance below is <= 2x the error MEASURED on B200 (profiles/r02_parity_errors.json keeps the measured value next to it).  Dense
conv stacks are compared relative to the activation scale max|x| (G2 activations reach several hundred): the absolute 1e-3 of SURVEY 8c is
met on everything that leaves the path (descriptors 2e-4, heat map 2e-7, line scores 8e-5, match probabilities 4e-4) while intermediate feature
maps carry the fp16 storage rounding of 70 stacked layers (measured <= 2.4e-3 of the scale at the deepest tap).

Discrete stages (NMS, top-k, association, unique pairs, line acceptance) are compared EXACTLY against the oracle
stage applied to the GPU's own upstream tensor; dense stages are compared with the precision-matched ("emul":
fp16 operands, fp32 accumulate) oracle within the tolerances written below.  End-to-end agreement with the pure
oracle is reported as set overlap (SURVEY.md §8c tolerance statement)."""
import numpy as np
import pytest
import torch

import _parity as P

pytestmark = pytest.mark.gpu

CFG = dict(max_keypoints=400, keypoint_threshold=0.004, remove_borders=4, line_threshold=0.75, line_length_threshold=50.0)


@pytest.fixture(scope="module")
def ctx():
    from airslam_b200 import capi
    c = capi.Context(max_batch=2, enable_lightglue=0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def images():
    from oracle import synth
    l, r, _ = synth.stereo_pair(752, 480, 21)
    return np.stack([l, r])


def _nhwc_to_nchw(a):
    return np.ascontiguousarray(a.transpose(2, 0, 1))


def _rel(a, b):
    """max |a - b| relative to the activation scale max(1, max |b|)."""
    return float(np.abs(a - b).max() / max(1.0, float(np.abs(b).max())))


def expected_junctions(adj, line_score, scores, desc_raw, border, in_w, in_h):
    """PLNet::process_output's junction_map marking (src/plnet.cpp:518-541) + junction_detector (:425-448) restated on the GPU's OWN
    stage-1 outputs (lines_adjusted, scores_line), heat map and dense descriptors: what the junction matrix must be, bit for bit."""
    from oracle import host
    jm = np.zeros((512, 512), dtype=bool)
    f32 = np.float32
    for k in range(len(line_score)):
        if line_score[k] < 0.5:
            continue
        x1, y1, x2, y2 = (f32(adj[k, c]) * f32(4) for c in range(4))
        xi1, yi1, xi2, yi2 = int(x1 + f32(0.1)), int(y1 + f32(0.1)), int(x2 + f32(0.1)), int(y2 + f32(0.1))
        jm[yi1, xi1] = (xi1 > border) and (xi1 < 512 - border) and (yi1 > border) and (yi1 < 512 - border)
        jm[yi2, xi2] = (xi2 > border) and (xi2 < 512 - border) and (yi2 > border) and (yi2 < 512 - border)
    sub = jm[border:512 - border, border:512 - border]
    yy, xx = np.nonzero(sub)
    yy, xx = yy + border, xx + border
    pts = np.stack([scores[yy, xx], xx.astype(f32), yy.astype(f32)]).astype(f32)
    d = desc_raw / np.maximum(np.sqrt((desc_raw * desc_raw).sum(-1, keepdims=True)), 1e-12)
    j = np.concatenate([pts, host.extract_descriptors(_nhwc_to_nchw(d), pts)], axis=0).astype(f32)
    j[1] *= f32(in_w) / f32(512)
    j[2] *= f32(in_h) / f32(512)
    return j


def assert_junctions(tag, got, exp):
    """exact count, (x, y) and order; scores exact; descriptors <= 1e-5."""
    P.exact(tag + ".count+xy+score", got.shape == exp.shape and np.array_equal(got[:3], exp[:3]))
    if exp.shape[1]:
        P.check(tag + ".descriptors", np.abs(got[3:] - exp[3:]).max(), 2e-7)


def test_superpoint_stages(ctx, images):
    from airslam_b200 import capi
    from oracle import host, nets, weights
    w = weights.load("superpoint")
    res = ctx.detect_batch(capi.NET_SUPERPOINT, images)
    for i in range(2):
        img = images[i]
        x = host.process_image(img)
        # K1 resize: bit exact (fp16 rounding of the reference float)
        x16 = ctx.debug_read(capi.NET_SUPERPOINT, "x16", i, np.float16, (512, 512))
        P.exact("K1.resize 752x480 -> 512x512 (+/255, fp16)", np.array_equal(x16, x[0, 0].astype(np.float16)))
        keep = {}
        sc_o, de_o = nets.superpoint_forward(x, w, emul=True, keep=keep)
        # conv stack: fp16 activations vs fp32-activation oracle with fp16 operand rounding
        # (relu_1, conv1b's full-resolution output, is only materialised by PLNet contexts -- its line branch reads it; checked in test_plnet_stages)
        r7 = ctx.debug_read(capi.NET_SUPERPOINT, "relu_7", i, np.float16, (64, 64, 128)).astype(np.float32)
        r7_o = keep["relu_7"][0].numpy().transpose(1, 2, 0)
        P.check("G1.relu_7 (8 convs)", _rel(r7, r7_o), 1e-3, "rel. to activation scale")
        logits = ctx.debug_read(capi.NET_SUPERPOINT, "logits", i, np.float32, (64, 64, 80))[..., :65]
        lg_o = keep["logits"][0].numpy().transpose(1, 2, 0)
        P.check("G1.logits (convPb)", _rel(logits, lg_o), 1e-3, "rel. to activation scale")
        draw = ctx.debug_read(capi.NET_SUPERPOINT, "desc_raw", i, np.float32, (64, 64, 256))
        dr_o = keep["desc_raw"][0].numpy().transpose(1, 2, 0)
        P.check("G1.desc_raw (convDb)", _rel(draw, dr_o), 1.6e-3, "rel. to activation scale")
        # K4 softmax + depth-to-space on OUR logits
        heat = ctx.debug_read(capi.NET_SUPERPOINT, "heat", i, np.float32, (512, 512))
        prob = torch.softmax(torch.from_numpy(logits), dim=-1)[..., :64].numpy()
        heat_o = prob.reshape(64, 64, 8, 8).transpose(0, 2, 1, 3).reshape(512, 512)
        P.check("K4.heat (softmax+d2s on own logits)", np.abs(heat - heat_o).max(), 4e-7)
        # K5 NMS on OUR heat: exact
        scores = ctx.debug_read(capi.NET_SUPERPOINT, "scores", i, np.float32, (512, 512))
        sc_ref = nets.simple_nms(torch.from_numpy(heat)[None])[0].numpy()
        P.exact("K5.simple_nms", np.array_equal(scores, sc_ref))
        # K6 keypoints on OUR scores: exact, including order
        feat = res[i][0]
        pts = host.detect_point(scores, CFG["keypoint_threshold"], CFG["remove_borders"], CFG["max_keypoints"])
        ws, hs = np.float32(752) / np.float32(512), np.float32(480) / np.float32(512)
        P.exact("K6.detect_point (count, order, score, xy)", feat.shape[1] == pts.shape[1] and np.array_equal(feat[0], pts[0]) and
                np.array_equal(feat[1], pts[1] * ws) and np.array_equal(feat[2], pts[2] * hs))
        # K7+K8 descriptors from OUR dense map at OUR keypoints: 1e-5 abs
        d = draw / np.maximum(np.sqrt((draw * draw).sum(-1, keepdims=True)), 1e-12)
        desc_ref = host.extract_descriptors(_nhwc_to_nchw(d), pts)
        P.check("K7+K8.sampled descriptors (own dense map)", np.abs(feat[3:] - desc_ref).max(), 2e-7)
        # end to end vs the pure oracle (emul mode): set overlap of keypoints, descriptor distance on the common ones
        f_o = host.keypoints_decoder(sc_o[0].numpy(), de_o[0].numpy(), CFG["keypoint_threshold"], CFG["remove_borders"], CFG["max_keypoints"])
        ours = {(int(a), int(b)): k for k, (a, b) in enumerate(zip(pts[1], pts[2]))}
        common = [(ours[(int(a), int(b))], k) for k, (a, b) in enumerate(zip(f_o[1], f_o[2])) if (int(a), int(b)) in ours]
        P.report("G1.e2e keypoint overlap vs emul oracle", len(common) / max(1, f_o.shape[1]), "fraction", "reported, not a gate beyond 0.95")
        assert len(common) >= 0.95 * f_o.shape[1], "keypoint overlap %d / %d" % (len(common), f_o.shape[1])
        dd = max(np.abs(feat[3:, a] - f_o[3:, b]).max() for a, b in common)
        P.check("G1.e2e descriptors on common keypoints", dd, 5e-4)


def test_plnet_stages(ctx, images):
    from airslam_b200 import capi
    from oracle import host, nets, weights
    w = weights.load("plnet")
    res = ctx.detect_batch(capi.NET_PLNET, images, lines=True, junctions=True)
    N = capi.NET_PLNET
    for i in range(2):
        img = images[i]
        x = host.process_image(img)
        keep = {}
        o = nets.plnet_s0_forward(x, w, emul=True, keep=keep)
        r1 = ctx.debug_read(N, "relu_1", i, np.float16, (512, 512, 64)).astype(np.float32)
        P.check("G2.relu_1 (conv1a+conv1b, fp16 store)", _rel(r1, keep["relu_1"][0].numpy().transpose(1, 2, 0)), 8e-4, "rel. to activation scale")
        heads9 = ctx.debug_read(N, "heads9", i, np.float32, (128, 128, 16))[..., :9]
        h_o = keep["heads9"][0].numpy().transpose(1, 2, 0)
        P.check("G2.heads9 (74 convs)", _rel(heads9, h_o), 5e-3, "rel. to activation scale")
        loi = ctx.debug_read(N, "loi", i, np.float32, (128, 128, 128))
        ta = ctx.debug_read(N, "thinaux", i, np.float32, (128, 128, 8))
        P.check("G2.loi_features", _rel(loi, o["loi_features"][0].numpy().transpose(1, 2, 0)), 2e-3, "rel. to activation scale")
        ta_o = np.concatenate([o["loi_features_thin"][0].numpy(), o["loi_features_aux"][0].numpy()]).transpose(1, 2, 0)
        P.check("G2.thin/aux", _rel(ta, ta_o), 1.5e-3, "rel. to activation scale")
        # K9 decode on OUR heads: lines within 1e-3 grid units; junction indices exact given our jloc
        dec = nets.hafm_decode(torch.from_numpy(np.ascontiguousarray(heads9.transpose(2, 0, 1)))[None])
        lines = ctx.debug_read(N, "lines_pred", i, np.float32, (3 * 128 * 128, 4))
        # tan() near pi/2 amplifies 1-ulp differences of sin/cos/tan between CUDA libm and the CPU: loose max, tight median
        dl = np.abs(lines - dec["lines_pred"].numpy())
        P.check("K9.lines_pred max (own heads; tan near pi/2)", dl.max(), 2e-3, "grid units")
        P.check("K9.lines_pred median", np.median(dl), 1e-5, "grid units")
        jloc = ctx.debug_read(N, "jloc", i, np.float32, (128, 128))
        P.check("K9.jloc", np.abs(jloc - dec["jloc"][0, 0].numpy()).max(), 3e-7)
        joff = dec["joff"]
        ja = nets.junctions_and_association(torch.from_numpy(lines), torch.from_numpy(jloc)[None, None], joff)
        jidx = ctx.debug_read(N, "junc_idx", i, np.int32, (300,))
        P.exact("K9.junction TopK indices", np.array_equal(jidx, ja["junc_topk_idx"].numpy().astype(np.int32)))
        juncs = ctx.debug_read(N, "juncs_pred", i, np.float32, (300, 2))
        P.check("K9.juncs_pred", np.abs(juncs - ja["juncs_pred"].numpy()).max(), 1e-5, "grid units")
        # K10 association on OUR lines + OUR junctions: exact integers
        ja2 = nets.junctions_and_association(torch.from_numpy(lines), torch.from_numpy(jloc)[None, None], joff)
        d1 = ((torch.from_numpy(lines)[None, :, 0:2] - torch.from_numpy(juncs)[:, None, :]) ** 2).sum(-1)
        d2 = ((torch.from_numpy(lines)[None, :, 2:4] - torch.from_numpy(juncs)[:, None, :]) ** 2).sum(-1)
        i1, i2 = np.argmin(d1.numpy(), 0), np.argmin(d2.numpy(), 0)
        imin_o, imax_o = np.minimum(i1, i2), np.maximum(i1, i2)
        keep_o = (imin_o < imax_o) & (d1.amin(0).numpy() < 10.0) & (d2.amin(0).numpy() < 10.0)
        imin = ctx.debug_read(N, "imin", i, np.int32, (49152,))
        imax = ctx.debug_read(N, "imax", i, np.int32, (49152,))
        iskeep = ctx.debug_read(N, "iskeep", i, np.uint8, (49152,))
        # the keep mask is exact for every proposal; (imin, imax) are defined -- and read downstream -- only where keep is set
        # (the device prunes the junction scan to the sqrt(10)-px neighbourhood, plnet.cpp:272-307 uses kept rows only)
        P.exact("K10.association keep mask (49152 proposals)", np.array_equal(iskeep.astype(bool), keep_o))
        P.exact("K10.association (imin, imax) on kept rows", np.array_equal(imin[keep_o], imin_o[keep_o]) and np.array_equal(imax[keep_o], imax_o[keep_o]))
        assert keep_o.sum() > 100
        # unique pairs (wireframe_matcher) on OUR association: exact, including order
        keep_idx, inverse, pairs = host.wireframe_matcher(iskeep.astype(np.float32), imin.astype(np.float32), imax.astype(np.float32))
        nu = int(ctx.debug_read(N, "n_unique", i, np.int32, (1,))[0])
        up = ctx.debug_read(N, "uid_pairs", i, np.int32, (16384, 2))[:nu]
        P.exact("wireframe_matcher unique pairs (order included)", nu == len(pairs) and np.array_equal(up, pairs.astype(np.int32)))
        first = np.full(nu, -1, dtype=np.int64)
        first[inverse[::-1]] = np.arange(len(inverse))[::-1]
        uf = ctx.debug_read(N, "uid_first", i, np.int32, (16384,))[:nu]
        P.exact("wireframe_matcher first proposal per pair", np.array_equal(uf, keep_idx[first].astype(np.int32)))
        # K11 stage 1 on OUR inputs (precision matched): scores within 2e-3
        kp2 = {}
        adj, sl = nets.plnet_s1_forward(juncs, lines, pairs.astype(np.float32), inverse.astype(np.float32), keep_idx.astype(np.float32),
                                        loi.transpose(2, 0, 1)[None], ta.transpose(2, 0, 1)[None, :4], ta.transpose(2, 0, 1)[None, 4:], w,
                                        emul=True, keep=kp2)
        f496 = ctx.debug_read(N, "feat496", i, np.float16, (16384, 512))[:nu, :496].astype(np.float32)
        P.check("K11.feat496 (LOI gather, fp16 store)", _rel(f496, kp2["feat"].numpy()), 6e-4, "rel. to activation scale")
        ls = ctx.debug_read(N, "line_score", i, np.float32, (16384,))[:nu]
        P.check("G3.scores_line (own inputs)", np.abs(ls - sl.numpy()).max(), 2e-4)
        # K12 acceptance on OUR scores: exact line list
        adj_g = ctx.debug_read(N, "lines_adjusted", i, np.float32, (16384, 4))[:nu]
        P.exact("G3.lines_adjusted", np.array_equal(adj_g, adj.numpy()))
        exp = []
        for k in range(nu):
            if ls[k] < 0.5 or ls[k] < np.float32(CFG["line_threshold"]):
                continue
            p = adj_g[k] * np.float32(4)
            l2 = (p[2] - p[0]) * (p[2] - p[0]) + (p[3] - p[1]) * (p[3] - p[1])
            if l2 < np.float32(2500.0):
                continue
            exp.append(p)
        exp = np.array(exp, dtype=np.float64).reshape(-1, 4)
        ws, hs = np.float64(np.float32(752) / np.float32(512)), np.float64(np.float32(480) / np.float32(512))
        exp *= np.array([ws, hs, ws, hs])
        got = res[i][1]
        P.exact("K12.accepted line list (order included)", got.shape == exp.shape and np.array_equal(got, exp))
        # K12b junction features (junction_map marking + junction_detector + second descriptor sampler) on OUR stage-1 outputs, heat map
        # and dense descriptors: exact count / order / (x, y) / score, descriptors 1e-5 (src/plnet.cpp:425-448, 518-541, 565-571)
        scores = ctx.debug_read(N, "scores", i, np.float32, (512, 512))
        draw = ctx.debug_read(N, "desc_raw", i, np.float32, (64, 64, 256))
        j_exp = expected_junctions(adj_g, ls, scores, draw, CFG["remove_borders"], 752, 480)
        assert j_exp.shape[1] > 10, "the synthetic frame must produce junctions for this check to mean anything"
        assert_junctions("K12b.junctions (detect_batch)", res[i][2], j_exp)
        # keypoints of the PLNet path: same stage-wise exactness as G1 (detect_point on OUR scores)
        pts = host.detect_point(scores, CFG["keypoint_threshold"], CFG["remove_borders"], CFG["max_keypoints"])
        f32 = np.float32
        P.exact("G2.keypoints (detect_point on own scores)", res[i][0].shape[1] == pts.shape[1] and np.array_equal(res[i][0][0], pts[0]) and
                np.array_equal(res[i][0][1], pts[1] * (f32(752) / f32(512))) and np.array_equal(res[i][0][2], pts[2] * (f32(480) / f32(512))))
        # end to end vs the pure oracle (emul): REPORTED as set overlaps -- the discrete stages above are the gates (one fp16
        # rounding flip upstream legitimately moves a score across a threshold)
        f_o, l_o, j_o = host.plnet_process_output(o, w, CFG, 752, 480, True, emul=True)
        key = lambda a: {tuple(np.round(r, 1)) for r in a}
        P.report("G2.e2e line overlap vs emul oracle", len(key(got) & key(l_o)) / max(1, len(l_o)), "fraction")
        kp = lambda f: {(round(float(a), 2), round(float(b), 2)) for a, b in zip(f[1], f[2])}
        P.report("G2.e2e keypoint overlap vs emul oracle", len(kp(res[i][0]) & kp(f_o)) / max(1, f_o.shape[1]), "fraction")
        P.report("G2.e2e junction overlap vs emul oracle", len(kp(res[i][2]) & kp(j_o)) / max(1, j_o.shape[1]), "fraction")
        assert len(key(got) & key(l_o)) >= 0.8 * max(1, len(l_o)) and len(kp(res[i][0]) & kp(f_o)) >= 0.95 * f_o.shape[1]


def test_stereo_entry_junctions_pinned_and_pageable(ctx, images):
    """airfe_detect_match_stereo_batch hands the left image's junction features back through two different D2H routes (straight into a
    pinned caller buffer, or through internal staging for a pageable one): both must equal airfe_detect_batch's, which
    test_plnet_stages pins against the oracle."""
    import ctypes as C
    from airslam_b200 import capi
    c2 = capi.Context(max_batch=1, enable_superpoint=0)
    try:
        det = c2.detect_batch(capi.NET_PLNET, images, lines=True, junctions=True)
        st = c2.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, images[0:1], images[1:2], lines=True, junctions=True)[0]   # pinned junc buffer
        P.exact("stereo entry junctions (pinned buffer) == detect_batch", np.array_equal(st["junc"], det[0][2]) and det[0][2].shape[1] > 10)
        P.exact("stereo entry features / lines == detect_batch", np.array_equal(st["feat_l"], det[0][0]) and np.array_equal(st["feat_r"], det[1][0]) and
                np.array_equal(st["lines_l"], det[0][1]) and np.array_equal(st["lines_r"], det[1][1]))
        # pageable buffers for everything
        L = capi.lib()
        fc, lc, jc, mc = 400, 2048, 512, 1024
        feat = np.zeros((2, fc, 259), np.float32); nf = np.zeros(2, np.int32)
        ln = np.zeros((2, lc, 4), np.float64); nl = np.zeros(2, np.int32)
        jn = np.zeros((1, jc, 259), np.float32); nj = np.zeros(1, np.int32)
        i0 = np.zeros((1, mc), np.int32); i1 = np.zeros((1, mc), np.int32); sc = np.zeros((1, mc), np.float32); nm = np.zeros(1, np.int32)
        q = lambda a: a.ctypes.data_as(C.c_void_p)
        l, r = np.ascontiguousarray(images[0]), np.ascontiguousarray(images[1])
        rc = L.airfe_detect_match_stereo_batch(c2.h, capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, 1, q(l), q(r), 752, 480, 752, 752 * 480, q(feat), fc, q(nf),
                                               q(ln), lc, q(nl), q(jn), jc, q(nj), q(i0), q(i1), q(sc), mc, q(nm))
        assert rc == 0
        P.exact("stereo entry junctions (pageable buffer) == detect_batch", np.array_equal(jn[0, :nj[0]].T, det[0][2]))
        # junc without lines is rejected instead of leaving n_junc unwritten
        rc = L.airfe_detect_match_stereo_batch(c2.h, capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, 1, q(l), q(r), 752, 480, 752, 752 * 480, q(feat), fc, q(nf),
                                               None, lc, None, q(jn), jc, q(nj), q(i0), q(i1), q(sc), mc, q(nm))
        assert rc == -1 and b"junction detection needs line detection" in L.airfe_last_error(c2.h)
    finally:
        c2.close()
