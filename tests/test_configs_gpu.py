"""BASELINE.json configs 3 and 4 as GPU parity cases (configs[1] is the bench workload and is covered by test_detect_gpu /
test_match_gpu; configs[0] is the CPU plumbing case covered by the oracle tests)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config3_superpoint_superglue_640x480_batch8():
    """synthetic 640x480 stereo batch = 8, SuperPoint + SuperGlue-indoor (the per-GPU shard of config 3)."""
    from airslam_b200 import capi
    from oracle import host, synth, weights
    ims = [synth.stereo_pair(640, 480, 0xA1750003 + k) for k in range(8)]
    left = np.stack([a for a, _, _ in ims])
    right = np.stack([b for _, b, _ in ims])
    ctx = capi.Context(max_batch=8, enable_plnet=0, enable_lightglue=0, enable_superglue=1, image_width=640, image_height=480)
    out = ctx.stereo_batch(capi.NET_SUPERPOINT, capi.MATCHER_SUPERGLUE, left, right)
    two = ctx.match_batch(capi.MATCHER_SUPERGLUE, [o["feat_l"] for o in out], [o["feat_r"] for o in out])
    ctx.close()
    wsg = weights.load("superglue_indoor")
    for p in range(8):
        assert np.array_equal(out[p]["matches"][0], two[p][0])                  # device-resident path == host round trip
        i0, i1 = out[p]["matches"][0][:, 0], out[p]["matches"][0][:, 1]
        disp = out[p]["feat_l"][1, i0] - out[p]["feat_r"][1, i1]
        assert len(i0) > 100 and abs(np.median(disp) - ims[p][2] * 1.0) < 1.5
    # oracle on the GPU's features of pair 0: identical SuperGlue matches
    m = host.matching_points(out[0]["feat_l"], out[0]["feat_r"], wsg, 1, 640, 480, emul=True)
    assert [tuple(r) for r in out[0]["matches"][0]] == [(a, b) for a, b, _ in m]


def test_config4_lowlight_1280x720_plnet_lightglue():
    """1280x720 low-light (OIVIO-shape) stereo, PLNet (max_keypoints 450, line_threshold 0.8: vo_oivio.yaml:3,6) + LightGlue."""
    from airslam_b200 import capi
    from oracle import host, nets, synth, weights
    l, r, d = synth.stereo_pair(1280, 720, 0xA1750004, low_light=True)
    ctx = capi.Context(max_batch=1, enable_superpoint=0, max_keypoints=450, line_threshold=0.8, image_width=1280, image_height=720)
    out = ctx.stereo_batch(capi.NET_PLNET, capi.MATCHER_LIGHTGLUE, l[None], r[None], lines=True, junctions=True)[0]
    x16 = ctx.debug_read(capi.NET_PLNET, "x16", 0, np.float16, (512, 512))
    scores = ctx.debug_read(capi.NET_PLNET, "scores", 0, np.float32, (512, 512))
    ctx.close()
    assert np.array_equal(x16, host.process_image(l)[0, 0].astype(np.float16))   # resize 1280x720 -> 512x512 bit exact
    pts = host.detect_point(scores, 0.004, 4, 450)
    ws, hs = np.float32(1280) / np.float32(512), np.float32(720) / np.float32(512)
    assert np.array_equal(out["feat_l"][1], pts[1] * ws) and np.array_equal(out["feat_l"][2], pts[2] * hs)
    # LightGlue on the GPU features == oracle on the same features (indices identical)
    m = host.matching_points(out["feat_l"], out["feat_r"], weights.load("lightglue"), 0, 1280, 720, emul=True)
    assert [tuple(q) for q in out["matches"][0]] == [(a, b) for a, b, _ in m]
    i0, i1 = out["matches"][0][:, 0], out["matches"][0][:, 1]
    if len(i0) > 30:
        assert abs(np.median(out["feat_l"][1, i0] - out["feat_r"][1, i1]) - d) < 2.0
    assert out["lines_l"].shape[1] == 4 and out["junc"].shape[0] == 259


def test_config5_relocalization_batch_single_rank():
    """Scaled-down config 5: keyframe map with planted query sources, device-resident keyframe cache (airfe_kf_put), every query
    LightGlue-matched against 3 candidates in batched launches that read both sides in place (airfe_reloc_match); world = 1 here (the
    2-rank exchange logic is covered on CPU by tests/test_shard_gloo.py and on NCCL by tests/test_reloc_nccl_gpu.py)."""
    import _parity as P
    from airslam_b200 import capi, reloc
    from oracle import synth
    n_kf, n_q = 24, 8
    kfs = [synth.keypoint_set(400, 752, 480, 0xA1750005 + k) for k in range(n_kf)]
    rs = np.random.RandomState(5)
    src = rs.permutation(n_kf)[:n_q]
    queries = [synth.keypoint_set(360, 752, 480, 900 + q, perturb_of=kfs[src[q]])[0] for q in range(n_q)]
    cand = np.zeros((n_q, 3), dtype=np.int64)
    for q in range(n_q):
        others = [k for k in rs.permutation(n_kf) if k != src[q]][:2]
        c = [src[q]] + others
        rs.shuffle(c)
        cand[q] = c
    ctx = capi.Context(max_batch=8, enable_superpoint=0, enable_plnet=0)
    reloc.upload_keyframes(ctx, kfs)
    best, cnt, table = reloc.relocalize(ctx, capi.MATCHER_LIGHTGLUE, queries, cand, n_kf)
    assert np.array_equal(best, src)
    assert cnt.min() > 250
    wrong = np.sort(table, axis=1)[:, :2]
    assert wrong.max() < 40           # unrelated keyframes produce (almost) no matches
    # the cached / in-place path == the host-buffer path (airfe_match_batch), job by job, bit for bit
    import torch
    qcap = 360
    qf = torch.zeros(n_q, qcap, 259)
    for i, f in enumerate(queries):
        qf[i, :f.shape[1]] = torch.from_numpy(np.ascontiguousarray(f.T))
    jq = [q for q in range(n_q) for _ in range(3)]
    jk = [int(cand[q, c]) for q in range(n_q) for c in range(3)]
    for where in ("host", "device"):
        t = qf.cuda() if where == "device" else qf
        counts, lists = ctx.reloc_match(capi.MATCHER_LIGHTGLUE, t.data_ptr(), [f.shape[1] for f in queries], qcap, jq, jk, want_matches=True)
        ref = ctx.match_batch(capi.MATCHER_LIGHTGLUE, [queries[q] for q in jq[:8]], [kfs[k] for k in jk[:8]])
        P.exact("reloc_match (%s queries, cached keyframes) == match_batch (host features)" % where,
                all(np.array_equal(lists[j][0], ref[j][0]) and np.array_equal(lists[j][1], ref[j][1]) for j in range(8)))
        assert np.array_equal(counts.reshape(n_q, 3), table)
    ctx.close()
