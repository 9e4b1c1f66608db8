"""CPU, authoring container only: the oracle restatement against a node-by-node execution of the reference's ONNX files
on a seed the golden fixtures do not use.  Skipped where /root/reference is absent (e.g. the GPU box)."""
import os

import numpy as np
import pytest

REF = "/root/reference/output/"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference model files not present on this machine")


def _load16(path):
    from tools import onnx_reader as R
    g = R.load(REF + path)
    for k in list(g.init):
        v = g.init[k]
        if v.dtype == np.float32 and v.ndim >= 2 and v.size > 16 and "onnx::Mul" not in k:
            g.init[k] = v.astype(np.float16).astype(np.float32)
    return g


def test_superpoint_graph():
    from oracle import host, nets, synth, weights
    from tools import onnx_interp as I
    x = host.process_image(synth.stereo_pair(640, 480, 99)[1])
    out = I.Interp(_load16("superpoint_v1_sim_int32.onnx")).run({"input": x})
    sc, de = nets.superpoint_forward(x, weights.load("superpoint"))
    assert np.array_equal(out["scores"].numpy() > 0, sc.numpy() > 0)
    assert np.abs(out["scores"].numpy() - sc.numpy()).max() <= 1e-6
    assert np.abs(out["descriptors"].numpy() - de.numpy()).max() <= 1e-6


def test_plnet_graphs_chain():
    from oracle import host, nets, synth, weights
    from tools import onnx_interp as I
    w = weights.load("plnet")
    x = host.process_image(synth.stereo_pair(752, 480, 98)[0])
    out = I.Interp(_load16("plnet_s0.onnx")).run({"input": x})
    o = nets.plnet_s0_forward(x, w)
    for k in ("iskeep", "idx_junc_to_end_min", "idx_junc_to_end_max", "lines_pred", "juncs_pred", "scores"):
        assert np.array_equal(out[k].numpy().reshape(-1), o[k].numpy().reshape(-1)), k
    for k in ("loi_features", "loi_features_thin", "loi_features_aux", "descriptors"):
        assert np.abs(out[k].numpy() - o[k].numpy()).max() <= 1e-6, k
    ki, inv, pairs = host.wireframe_matcher(out["iskeep"].numpy(), out["idx_junc_to_end_min"].numpy(), out["idx_junc_to_end_max"].numpy())
    feeds = dict(juncs_pred=out["juncs_pred"].numpy(), lines_pred=out["lines_pred"].numpy(), idx_lines_for_junctions=pairs.astype(np.float32),
                 inverse=inv.astype(np.float32).reshape(-1, 1), iskeep_index=ki.astype(np.float32).reshape(-1, 1), loi_features=out["loi_features"].numpy(),
                 loi_features_thin=out["loi_features_thin"].numpy(), loi_features_aux=out["loi_features_aux"].numpy())
    o3 = I.Interp(_load16("plnet_s1.onnx")).run(feeds)
    adj, sl = nets.plnet_s1_forward(feeds["juncs_pred"], feeds["lines_pred"], feeds["idx_lines_for_junctions"], feeds["inverse"], feeds["iskeep_index"],
                                    feeds["loi_features"], feeds["loi_features_thin"], feeds["loi_features_aux"], w)
    assert np.array_equal(o3["lines_adjusted"].numpy(), adj.numpy())
    assert np.abs(o3["scores_line"].numpy() - sl.numpy()).max() <= 2e-6


def test_matcher_graphs():
    from oracle import host, nets, synth, weights
    from tools import onnx_interp as I
    f0 = synth.keypoint_set(96, 752, 480, 17)
    f1, _ = synth.keypoint_set(80, 752, 480, 18, perturb_of=f0)
    a, b = host.normalize_keypoints(f0, 752, 480, 0.5), host.normalize_keypoints(f1, 752, 480, 0.5)
    feeds = {"keypoints_0": a[1:3].T.copy()[None], "keypoints_1": b[1:3].T.copy()[None], "descriptors_0": a[3:].T.copy()[None], "descriptors_1": b[3:].T.copy()[None]}
    s = I.Interp(_load16("superpoint_lightglue.onnx")).run(feeds)["scores"][0].numpy()
    so = nets.lightglue_forward(a[1:3].T.copy(), b[1:3].T.copy(), a[3:].T.copy(), b[3:].T.copy(), weights.load("lightglue")).numpy()
    assert np.abs(np.exp(s) - np.exp(so)).max() <= 1e-4
    a, b = host.normalize_keypoints(f0, 752, 480, 0.7), host.normalize_keypoints(f1, 752, 480, 0.7)
    feeds = {"keypoints_0": a[1:3].T.copy()[None], "scores_0": a[0][None].copy(), "descriptors_0": a[3:][None].copy(),
             "keypoints_1": b[1:3].T.copy()[None], "scores_1": b[0][None].copy(), "descriptors_1": b[3:][None].copy()}
    z = I.Interp(_load16("superglue_indoor_sim_int32.onnx")).run(feeds)["scores"][0].numpy()
    zo = nets.superglue_forward(a[1:3].T.copy(), a[0].copy(), a[3:].copy(), b[1:3].T.copy(), b[0].copy(), b[3:].copy(), weights.load("superglue_indoor")).numpy()
    assert np.abs(z - zo).max() <= 1e-3
