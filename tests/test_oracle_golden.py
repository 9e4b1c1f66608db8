"""CPU: the oracle restatement reproduces the golden vectors frozen from the reference's own ONNX graphs
(tools/make_golden.py).  This is what pins the oracle on machines where /root/reference does not exist."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import host, nets, synth, weights

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def image():
    l, r, d = synth.stereo_pair(752, 480, 0xA175)
    return host.process_image(l)


def test_g1_superpoint(image):
    g = np.load(os.path.join(G, "g1_superpoint.npz"))
    sc, de = nets.superpoint_forward(image, weights.load("superpoint"))
    sc, de = sc[0].numpy(), de[0].numpy()
    nz = np.nonzero(sc.reshape(-1))[0]
    assert np.array_equal(nz, g["nz_idx"])
    assert np.abs(sc.reshape(-1)[nz] - g["nz_val"]).max() <= 1e-6
    assert np.abs(de[:, ::8, ::8] - g["desc_cells"]).max() <= 1e-6
    assert "drift_vs_fp32_weights" in json.loads(str(g["meta"]))


def test_g2_g3_plnet(image):
    g2 = np.load(os.path.join(G, "g2_plnet_s0.npz"))
    g3 = np.load(os.path.join(G, "g3_plnet_s1.npz"))
    w = weights.load("plnet")
    o = nets.plnet_s0_forward(image, w)
    assert np.abs(o["juncs_pred"].numpy() - g2["juncs_pred"]).max() <= 1e-5
    assert np.abs(o["lines_pred"].numpy()[::97] - g2["lines_pred_sub"]).max() <= 1e-4
    keep = np.nonzero(o["iskeep"].numpy() > 0)[0]
    assert np.array_equal(keep, g2["iskeep_idx"])
    assert np.array_equal(o["idx_junc_to_end_min"].numpy()[keep].astype(np.int16), g2["idx_min_kept"])
    assert np.array_equal(o["idx_junc_to_end_max"].numpy()[keep].astype(np.int16), g2["idx_max_kept"])
    assert np.abs(o["loi_features"].numpy()[0, :, ::16, ::16] - g2["loi_cells"]).max() <= 1e-5
    assert np.abs(o["loi_features_thin"].numpy()[0, :, ::4, ::4] - g2["thin"]).max() <= 1e-5
    assert np.array_equal(np.nonzero(o["scores"].numpy().reshape(-1))[0], g2["scores_nz"])
    ki, inv, pairs = host.wireframe_matcher(o["iskeep"].numpy(), o["idx_junc_to_end_min"].numpy(), o["idx_junc_to_end_max"].numpy())
    assert np.array_equal(pairs.astype(np.int16), g3["pairs"])
    adj, sl = nets.plnet_s1_forward(o["juncs_pred"], o["lines_pred"], pairs.astype(np.float32), inv.astype(np.float32), ki.astype(np.float32),
                                    o["loi_features"], o["loi_features_thin"], o["loi_features_aux"], w)
    assert np.abs(adj.numpy() - g3["lines_adjusted"]).max() <= 1e-5
    assert np.abs(sl.numpy() - g3["scores_line"]).max() <= 1e-5
    e = np.load(os.path.join(G, "plnet_e2e.npz"))
    feats, lines, junc = host.plnet_process_output(o, w, host.PLNET_CFG_EUROC, 752, 480, True)
    assert np.array_equal(feats[:3], e["feat_xy"]) or np.abs(feats[:3] - e["feat_xy"]).max() <= 1e-6
    assert np.abs(feats[:, :64] - e["feat_head"]).max() <= 1e-5
    assert lines.shape == e["lines"].shape and np.abs(lines - e["lines"]).max() <= 1e-4
    assert np.array_equal(junc[1:3], e["junc_xy"][1:3])


def _match_inputs(scale):
    f0 = synth.keypoint_set(160, 752, 480, 7)
    f1, perm = synth.keypoint_set(144, 752, 480, 8, perturb_of=f0)
    return host.normalize_keypoints(f0, 752, 480, scale), host.normalize_keypoints(f1, 752, 480, scale)


def test_g4_lightglue():
    g = np.load(os.path.join(G, "g4_lightglue.npz"))
    n0, n1 = _match_inputs(0.5)
    idx, sc, dense = host.lightglue_infer(n0[1:], n1[1:], weights.load("lightglue"))
    big = g["scores"] > np.log(1e-4)
    assert np.abs(np.exp(dense[big]) - np.exp(g["scores"][big])).max() <= 1e-4
    assert np.array_equal(idx, g["matches"])
    assert np.abs(sc - g["match_scores"]).max() <= 1e-4


@pytest.mark.parametrize("kind", ["indoor", "outdoor"])
def test_g5_superglue(kind):
    path = os.path.join(weights.WEIGHT_DIR, "superglue_%s.afw" % kind)
    if not os.path.exists(path):
        pytest.skip("weights for %s not shipped to this box" % kind)
    g = np.load(os.path.join(G, "g5_superglue_%s.npz" % kind))
    n0, n1 = _match_inputs(0.7)
    i0, i1, m0, m1, dense = host.superglue_infer(n0, n1, weights.load("superglue_" + kind))
    assert np.abs(dense - g["scores"]).max() <= 2e-3
    assert np.array_equal(i0, g["indices0"]) and np.array_equal(i1, g["indices1"])
    assert np.abs(m0 - g["mscores0"]).max() <= 1e-4
