"""Stage-wise + end-to-end GPU parity of the detector (SuperPoint G1, PLNet G2+G3) against the oracle.

Discrete stages (NMS, top-k, association, unique pairs, line acceptance) are compared EXACTLY against the oracle
stage applied to the GPU's own upstream tensor; dense stages are compared with the precision-matched ("emul":
fp16 operands, fp32 accumulate) oracle within the tolerances written below.  End-to-end agreement with the pure
oracle is reported as set overlap (SURVEY.md §8c tolerance statement)."""
import numpy as np
import pytest
import torch

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
        assert np.array_equal(x16, x[0, 0].astype(np.float16))
        keep = {}
        sc_o, de_o = nets.superpoint_forward(x, w, emul=True, keep=keep)
        # conv stack: fp16 activations vs fp32-activation oracle with fp16 operand rounding
        r1 = ctx.debug_read(capi.NET_SUPERPOINT, "relu_1", i, np.float16, (512, 512, 64)).astype(np.float32)
        r1_o = keep["relu_1"][0].numpy().transpose(1, 2, 0)
        assert np.abs(r1 - r1_o).max() <= 2e-3 * max(1.0, np.abs(r1_o).max())
        r7 = ctx.debug_read(capi.NET_SUPERPOINT, "relu_7", i, np.float16, (64, 64, 128)).astype(np.float32)
        r7_o = keep["relu_7"][0].numpy().transpose(1, 2, 0)
        assert np.abs(r7 - r7_o).max() <= 5e-3 * max(1.0, np.abs(r7_o).max())
        logits = ctx.debug_read(capi.NET_SUPERPOINT, "logits", i, np.float32, (64, 64, 80))[..., :65]
        lg_o = keep["logits"][0].numpy().transpose(1, 2, 0)
        assert np.abs(logits - lg_o).max() <= 5e-3 * max(1.0, np.abs(lg_o).max())
        draw = ctx.debug_read(capi.NET_SUPERPOINT, "desc_raw", i, np.float32, (64, 64, 256))
        dr_o = keep["desc_raw"][0].numpy().transpose(1, 2, 0)
        assert np.abs(draw - dr_o).max() <= 5e-3 * max(1.0, np.abs(dr_o).max())
        # K4 softmax + depth-to-space on OUR logits
        heat = ctx.debug_read(capi.NET_SUPERPOINT, "heat", i, np.float32, (512, 512))
        prob = torch.softmax(torch.from_numpy(logits), dim=-1)[..., :64].numpy()
        heat_o = prob.reshape(64, 64, 8, 8).transpose(0, 2, 1, 3).reshape(512, 512)
        assert np.abs(heat - heat_o).max() <= 1e-6
        # K5 NMS on OUR heat: exact
        scores = ctx.debug_read(capi.NET_SUPERPOINT, "scores", i, np.float32, (512, 512))
        sc_ref = nets.simple_nms(torch.from_numpy(heat)[None])[0].numpy()
        assert np.array_equal(scores, sc_ref)
        # K6 keypoints on OUR scores: exact, including order
        feat = res[i][0]
        pts = host.detect_point(scores, CFG["keypoint_threshold"], CFG["remove_borders"], CFG["max_keypoints"])
        assert feat.shape[1] == pts.shape[1]
        ws, hs = np.float32(752) / np.float32(512), np.float32(480) / np.float32(512)
        assert np.array_equal(feat[0], pts[0])
        assert np.array_equal(feat[1], pts[1] * ws) and np.array_equal(feat[2], pts[2] * hs)
        # K7+K8 descriptors from OUR dense map at OUR keypoints: 1e-5 abs
        d = draw / np.maximum(np.sqrt((draw * draw).sum(-1, keepdims=True)), 1e-12)
        desc_ref = host.extract_descriptors(_nhwc_to_nchw(d), pts)
        assert np.abs(feat[3:] - desc_ref).max() <= 1e-5
        # end to end vs the pure oracle (emul mode): set overlap of keypoints, descriptor distance on the common ones
        f_o = host.keypoints_decoder(sc_o[0].numpy(), de_o[0].numpy(), CFG["keypoint_threshold"], CFG["remove_borders"], CFG["max_keypoints"])
        ours = {(int(a), int(b)): k for k, (a, b) in enumerate(zip(pts[1], pts[2]))}
        common = [(ours[(int(a), int(b))], k) for k, (a, b) in enumerate(zip(f_o[1], f_o[2])) if (int(a), int(b)) in ours]
        assert len(common) >= 0.95 * f_o.shape[1], "keypoint overlap %d / %d" % (len(common), f_o.shape[1])
        dd = max(np.abs(feat[3:, a] - f_o[3:, b]).max() for a, b in common)
        assert dd <= 3e-3, dd


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
        heads9 = ctx.debug_read(N, "heads9", i, np.float32, (128, 128, 16))[..., :9]
        h_o = keep["heads9"][0].numpy().transpose(1, 2, 0)
        assert np.abs(heads9 - h_o).max() <= 2e-2 * max(1.0, np.abs(h_o).max()), np.abs(heads9 - h_o).max()
        loi = ctx.debug_read(N, "loi", i, np.float32, (128, 128, 128))
        assert np.abs(loi - o["loi_features"][0].numpy().transpose(1, 2, 0)).max() <= 2e-2 * max(1.0, float(o["loi_features"].abs().max()))
        ta = ctx.debug_read(N, "thinaux", i, np.float32, (128, 128, 8))
        ta_o = np.concatenate([o["loi_features_thin"][0].numpy(), o["loi_features_aux"][0].numpy()]).transpose(1, 2, 0)
        assert np.abs(ta - ta_o).max() <= 2e-2 * max(1.0, np.abs(ta_o).max())
        # K9 decode on OUR heads: lines within 1e-3 grid units; junction indices exact given our jloc
        dec = nets.hafm_decode(torch.from_numpy(np.ascontiguousarray(heads9.transpose(2, 0, 1)))[None])
        lines = ctx.debug_read(N, "lines_pred", i, np.float32, (3 * 128 * 128, 4))
        # tan() near pi/2 amplifies 1-ulp differences of sin/cos/tan between CUDA libm and the CPU: loose max, tight median
        dl = np.abs(lines - dec["lines_pred"].numpy())
        assert dl.max() <= 2e-2 and np.median(dl) <= 1e-5, (dl.max(), np.median(dl))
        jloc = ctx.debug_read(N, "jloc", i, np.float32, (128, 128))
        assert np.abs(jloc - dec["jloc"][0, 0].numpy()).max() <= 1e-6
        joff = dec["joff"]
        ja = nets.junctions_and_association(torch.from_numpy(lines), torch.from_numpy(jloc)[None, None], joff)
        jidx = ctx.debug_read(N, "junc_idx", i, np.int32, (300,))
        assert np.array_equal(jidx, ja["junc_topk_idx"].numpy().astype(np.int32))
        juncs = ctx.debug_read(N, "juncs_pred", i, np.float32, (300, 2))
        assert np.abs(juncs - ja["juncs_pred"].numpy()).max() <= 1e-5
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
        assert np.array_equal(iskeep.astype(bool), keep_o)
        assert np.array_equal(imin[keep_o], imin_o[keep_o]) and np.array_equal(imax[keep_o], imax_o[keep_o])
        assert keep_o.sum() > 100
        # unique pairs (wireframe_matcher) on OUR association: exact, including order
        keep_idx, inverse, pairs = host.wireframe_matcher(iskeep.astype(np.float32), imin.astype(np.float32), imax.astype(np.float32))
        nu = int(ctx.debug_read(N, "n_unique", i, np.int32, (1,))[0])
        assert nu == len(pairs)
        up = ctx.debug_read(N, "uid_pairs", i, np.int32, (16384, 2))[:nu]
        assert np.array_equal(up, pairs.astype(np.int32))
        first = np.full(nu, -1, dtype=np.int64)
        first[inverse[::-1]] = np.arange(len(inverse))[::-1]
        uf = ctx.debug_read(N, "uid_first", i, np.int32, (16384,))[:nu]
        assert np.array_equal(uf, keep_idx[first].astype(np.int32))
        # K11 stage 1 on OUR inputs (precision matched): scores within 2e-3
        kp2 = {}
        adj, sl = nets.plnet_s1_forward(juncs, lines, pairs.astype(np.float32), inverse.astype(np.float32), keep_idx.astype(np.float32),
                                        loi.transpose(2, 0, 1)[None], ta.transpose(2, 0, 1)[None, :4], ta.transpose(2, 0, 1)[None, 4:], w,
                                        emul=True, keep=kp2)
        f496 = ctx.debug_read(N, "feat496", i, np.float16, (16384, 512))[:nu, :496].astype(np.float32)
        assert np.abs(f496 - kp2["feat"].numpy()).max() <= 4e-3 * max(1.0, float(kp2["feat"].abs().max()))
        ls = ctx.debug_read(N, "line_score", i, np.float32, (16384,))[:nu]
        assert np.abs(ls - sl.numpy()).max() <= 5e-3, np.abs(ls - sl.numpy()).max()
        # K12 acceptance on OUR scores: exact line list
        adj_g = ctx.debug_read(N, "lines_adjusted", i, np.float32, (16384, 4))[:nu]
        assert np.array_equal(adj_g, adj.numpy())
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
        assert got.shape == exp.shape and np.array_equal(got, exp), (got.shape, exp.shape)
        # end to end vs pure oracle: lines as a set (rounded), junction count
        f_o, l_o, j_o = host.plnet_process_output(o, w, CFG, 752, 480, True, emul=True)
        key = lambda a: {tuple(np.round(r, 1)) for r in a}
        inter = len(key(got) & key(l_o))
        assert inter >= 0.8 * max(1, len(l_o)), "line overlap %d / %d (ours %d)" % (inter, len(l_o), len(got))
        assert res[i][0].shape[1] == f_o.shape[1] or abs(res[i][0].shape[1] - f_o.shape[1]) <= 8
        assert res[i][2].shape[0] == 259
