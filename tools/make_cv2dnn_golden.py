"""EXTERNAL pin of the oracle: the reference's own ONNX files executed by OpenCV DNN (cv2.dnn 4.13) -- an independent runtime.

    python -m tools.make_cv2dnn_golden          (authoring container only: needs /root/reference and cv2)

The reference's arithmetic for this path lives in TensorRT 8.6 (absent; no sm_100 support) and the reference holds no golden
vectors, so round 1 pinned the oracle against tools/onnx_interp.py -- a second builder-written executor.  This
script replaces that by a third-party one.  cv2.dnn rejects the five shipped graphs as a whole (dynamic H x W, ConstantOfShape,
logical ops in the in-graph NMS / HAFM decode), so tools/onnx_cut.py re-emits static-shape sub-graphs with the nodes and
initialisers copied as raw bytes from the reference files:

  G1 superpoint_v1      input -> softmax(semi) ('49': 65 x H/8 x W/8) and 'descriptors' (L2-normalised)      [whole dense part]
  G2 plnet_s0           input -> fc2, the five head maps (Concat of heads.*.2), loi_features / _thin / _aux, point-detector
                        logits and raw descriptors                                                            [whole dense part]
  G3 plnet_s1           the verification MLP: 496-d line features -> fc2.* / fc2_res / fc2_head logits
  G4 superpoint_lightglue  WHOLE graph (keypoints / descriptors -> log-assignment scores)
  G1 nms                the in-graph simple_nms, arithmetic nodes step-wise in cv2.dnn, boolean And / Or / Not in numpy -> maximum mask
  G2 decode             head maps -> HAFM line decode ('lines_pred') and junction-heat NMS ('/Mul_17_output_0')
  G2 association        squared distances line end <-> junction + column minima in cv2.dnn, arg-min / Less / And in numpy -> idx_min / idx_max / iskeep
  G5 superglue_indoor   keypoint encoder + 18 GNN layers + final_proj + score einsum / sqrt(256) -> similarity matrix ('2435');
                        the 100 Sinkhorn iterations ('2563' .. 'scores') on the couplings built from it -> final score matrix

What cv2.dnn cannot execute is exactly the integer / logical tail of G1 / G2 (3-round NMS via MaxPool+Equal+Where, HAFM decode, TopK,
association): those are checked bit-exactly against tools/onnx_interp.py (tests/test_oracle_golden.py) and are integer arithmetic.

Weights: fp32 initialisers rounded to fp16-representable values (what the reference's kFP16 engines hold and weights/*.afw ship), the
same rule as tools/make_golden.py; re-encoded as raw fp32 data.  Outputs are stored subsampled (+ full-tensor max / mean for scale).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import host, synth  # noqa: E402
from tools import onnx_cut, onnx_reader as R  # noqa: E402

REF = "/root/reference/output/"
OUT = os.environ.get("AIRFE_GOLDEN_OUT", os.path.join(ROOT, "tests", "golden"))
TMP = "/tmp/airfe_cv2dnn"


def run(src, inputs, outputs, feeds, int_inputs=()):
    import cv2
    os.makedirs(TMP, exist_ok=True)
    dst = os.path.join(TMP, os.path.basename(src))
    onnx_cut.cut(REF + src, dst, inputs, {o: None for o in outputs}, int_inputs, fp16_weights=True)
    net = cv2.dnn.readNetFromONNX(dst)
    for k, v in feeds.items():
        net.setInput(np.ascontiguousarray(v, dtype=np.float32), k)
    names = list(net.getUnconnectedOutLayersNames())
    res = net.forward(names)
    return {n: np.asarray(r) for n, r in zip(names, res)}


def g1_inputs():
    l, _, _ = synth.stereo_pair(752, 480, 0xA175)
    return host.process_image(l)             # [1,1,480,752] float32 in [0,1]


def lg_inputs():
    f0 = synth.keypoint_set(160, 752, 480, 7)
    f1, _ = synth.keypoint_set(144, 752, 480, 8, perturb_of=f0)
    return f0, f1


def main():
    import cv2
    os.makedirs(OUT, exist_ok=True)
    meta = {"runtime": "cv2.dnn " + cv2.__version__, "weights": "fp32 initialisers rounded to fp16 values", "tool": "tools/make_cv2dnn_golden.py"}

    # ---- G1: SuperPoint dense part on the 512x512 network input (src/super_point.cpp:111-116)
    x = g1_inputs()
    o = run("superpoint_v1_sim_int32.onnx", {"input": list(x.shape)}, ["49", "descriptors"], {"input": x})
    np.savez_compressed(os.path.join(OUT, "cv2dnn_g1_superpoint.npz"), prob=o["49"][0].astype(np.float32), desc=o["descriptors"][0][:, ::2, ::2].copy(),
                        meta=json.dumps(dict(meta, image="oracle.synth.stereo_pair(752,480,0xA175)[0] through host.process_image")))
    print("G1", o["49"].shape, o["descriptors"].shape)
    # The in-graph NMS behind it (nodes '86'..'scores': five 9x9 MaxPools, Equal, Cast, Greater, Where, And / Or / Not) is executed STEP-WISE: cv2.dnn has
    # no And / Or / Not layers, so every arithmetic node group (MaxPool + Equal + Cast; MaxPool + Greater + Where + MaxPool) runs in cv2.dnn on sub-graphs
    # cut at the boolean nodes, and only the three boolean combinations per round are done here in numpy.  The heat map is the depth-to-space of
    # cv2.dnn's own softmax output (Slice / Reshape / Transpose glue, nodes '54'..'83').
    p64 = o["49"][0].astype(np.float32)[:64]                                                                 # drop the dustbin channel
    heat = np.ascontiguousarray(p64.transpose(1, 2, 0).reshape(64, 64, 8, 8).transpose(0, 2, 1, 3).reshape(1, 512, 512))
    shp, zeros = list(heat.shape), np.zeros_like(heat)
    m = run("superpoint_v1_sim_int32.onnx", {"83": shp}, ["90"], {"83": heat})["90"].reshape(heat.shape) > 0.5        # M0 = (S == mp9(S))
    for mask_in, mp_mask, s_out, mp_out in (("90", "93", "96", "99"), ("110", "113", "116", "119")):
        r = run("superpoint_v1_sim_int32.onnx", {mask_in: shp, "83": shp, "85": shp}, [mp_mask, s_out, mp_out],
                {mask_in: m.astype(np.float32), "83": heat, "85": zeros})
        supp = r[mp_mask].reshape(heat.shape) > 0                                                            # Greater('93', 0)
        m = m | ((r[s_out].reshape(heat.shape) == r[mp_out].reshape(heat.shape)) & ~supp)                    # Equal, Not, And, Or
    np.savez_compressed(os.path.join(OUT, "cv2dnn_g1_nms.npz"), mask_bits=np.packbits(m[0]), n_maxima=np.int64(m.sum()),
                        meta=json.dumps(dict(meta, what="NMS maximum mask [512][512] (np.packbits) of the heat map built from cv2dnn_g1_superpoint.npz 'prob'; scores = where(mask, heat, 0)")))
    print("G1 nms", int(m.sum()))

    # ---- G2: PLNet stage 0 dense part (512x512 network input as the reference resizes to, src/plnet.cpp:246-270)
    x2 = x
    heads = "/backbone/score1/Concat_output_0"
    pd_logits = "/backbone/point_detector/convPb/Conv_output_0"
    pd_desc = "/backbone/point_detector/convDb/Conv_output_0"
    outs = ["/backbone/fc2/Conv_output_0", heads, "loi_features", "loi_features_thin", "loi_features_aux", pd_logits, pd_desc]
    o = run("plnet_s0.onnx", {"input": list(x2.shape)}, outs, {"input": x2})
    np.savez_compressed(os.path.join(OUT, "cv2dnn_g2_plnet_s0.npz"), fc2=o[outs[0]][0][:, ::8, ::8].copy(), heads9=o[heads][0].astype(np.float32),
                        loi=o["loi_features"][0][:, ::8, ::8].copy(), thin=o["loi_features_thin"][0][:, ::2, ::2].copy(),
                        aux=o["loi_features_aux"][0][:, ::2, ::2].copy(), pd_logits=o[pd_logits][0].astype(np.float32),
                        pd_desc=o[pd_desc][0][:, ::4, ::4].copy(), meta=json.dumps(dict(meta, image="same frame, resized to 512x512 by the restated cv::resize")))
    print("G2", {k: v.shape for k, v in o.items()})
    # The in-graph HAFM decode behind the head maps (157 nodes: Sigmoid, Cos / Sin / Tan, Clip, the Range / Expand pixel grids -> 'lines_pred') and the
    # junction-heat NMS (Softmax, 3x3 MaxPool, Equal, Mul -> '/Mul_17_output_0') run in cv2.dnn when cut at the head tensor; the TopK / GatherElements
    # junction selection and the association behind them do not (cv2.dnn 4.13 crashes in TopK) and stay pinned by tools/onnx_interp.py.
    h9 = np.ascontiguousarray(o[heads].astype(np.float32))
    o_dec = run("plnet_s0.onnx", {heads: list(h9.shape)}, ["lines_pred", "/Mul_17_output_0"], {heads: h9})
    np.savez_compressed(os.path.join(OUT, "cv2dnn_g2_hafm_decode.npz"), lines_pred=o_dec["lines_pred"].astype(np.float32)[::4].copy(),
                        jloc_nms=o_dec["/Mul_17_output_0"].astype(np.float32).reshape(128, 128),
                        meta=json.dumps(dict(meta, inputs="heads9 of cv2dnn_g2_plnet_s0.npz (cv2.dnn's own head maps of the same frame); lines_pred rows ::4")))
    print("G2 decode", {k: v.shape for k, v in o_dec.items()})
    # Association (nodes 'Slice_10' .. 'iskeep'): the 300 x 49 152 squared-distance matrices of both line ends and their column minima run in cv2.dnn
    # (Sub, Pow, ReduceSum, ReduceMin), cut at (lines_pred, juncs_pred); cv2.dnn 4.13 returns garbage from the ArgMin -> Min / Max -> Cast chain, so the
    # arg-min (first index on ties), Min / Max / Less / And are taken here in numpy ON cv2.dnn's matrices.  juncs_pred comes from the oracle's TopK
    # (interpreter-pinned) applied to the same head maps; lines_pred is cv2.dnn's own.
    import torch
    from oracle import nets as _nets
    dec = _nets.hafm_decode(torch.from_numpy(h9))
    lp, jp = np.ascontiguousarray(o_dec["lines_pred"].astype(np.float32)), np.ascontiguousarray(dec["juncs_pred"].numpy())
    aouts = ["/ReduceSum_output_0", "/ReduceSum_1_output_0", "/ReduceMin_output_0", "/ReduceMin_1_output_0"]
    oa = run("plnet_s0.onnx", {"lines_pred": list(lp.shape), "juncs_pred": list(jp.shape)}, aouts, {"lines_pred": lp, "juncs_pred": jp})
    d1, d2 = oa[aouts[0]].reshape(300, -1), oa[aouts[1]].reshape(300, -1)
    m1, m2 = oa[aouts[2]].reshape(-1), oa[aouts[3]].reshape(-1)
    i1, i2 = np.argmin(d1, 0), np.argmin(d2, 0)
    imin, imax = np.minimum(i1, i2), np.maximum(i1, i2)
    keep = (imin < imax) & (m1 < 10.0) & (m2 < 10.0)
    np.savez_compressed(os.path.join(OUT, "cv2dnn_g2_association.npz"), juncs_pred=jp, idx_min=imin.astype(np.uint16), idx_max=imax.astype(np.uint16),
                        iskeep_bits=np.packbits(keep), n_keep=np.int64(keep.sum()),
                        meta=json.dumps(dict(meta, inputs="lines_pred = cv2.dnn's HAFM decode (all 49 152 rows), juncs_pred = oracle TopK on the same head maps")))
    print("G2 association", int(keep.sum()))

    # ---- G3: PLNet stage 1 verification MLP on the oracle's own 496-d line features of that frame (the sampler in front of it is
    # Gather / Floor / Clip index arithmetic: checked against tools/onnx_interp.py in tests/test_oracle_golden.py)
    import torch
    from oracle import nets, weights
    torch.set_num_threads(8)
    wpl = weights.load("plnet")
    g2 = nets.plnet_s0_forward(x, wpl)
    keep_idx, inverse, pairs = host.wireframe_matcher(g2["iskeep"].numpy(), g2["idx_junc_to_end_min"].numpy(), g2["idx_junc_to_end_max"].numpy())
    k3 = {}
    nets.plnet_s1_forward(g2["juncs_pred"].numpy(), g2["lines_pred"].numpy(), pairs.astype(np.float32), inverse.astype(np.float32).reshape(-1, 1),
                          keep_idx.astype(np.float32).reshape(-1, 1), g2["loi_features"].numpy(), g2["loi_features_thin"].numpy(),
                          g2["loi_features_aux"].numpy(), wpl, keep=k3)
    feat = k3["feat"].numpy().astype(np.float32)[:512]
    o = run("plnet_s1.onnx", {"/Concat_38_output_0": list(feat.shape), "/Concat_39_output_0": [feat.shape[0], 240]}, ["/fc2_head/Gemm_output_0"],
            {"/Concat_38_output_0": feat, "/Concat_39_output_0": feat[:, 256:].copy()})
    np.savez_compressed(os.path.join(OUT, "cv2dnn_g3_plnet_s1_mlp.npz"), feat=feat, logits=o["/fc2_head/Gemm_output_0"].astype(np.float32),
                        meta=json.dumps(dict(meta, inputs="first 512 rows of the oracle's 496-d line features of the G2 frame")))
    print("G3", o["/fc2_head/Gemm_output_0"].shape)

    # ---- G4: LightGlue, the WHOLE graph (inputs as tools/make_golden.py)
    f0, f1 = lg_inputs()
    n0, n1 = host.normalize_keypoints(f0, 752, 480, 0.5), host.normalize_keypoints(f1, 752, 480, 0.5)
    feeds = {"keypoints_0": n0[1:3].T.copy()[None], "keypoints_1": n1[1:3].T.copy()[None], "descriptors_0": n0[3:].T.copy()[None], "descriptors_1": n1[3:].T.copy()[None]}
    o = run("superpoint_lightglue.onnx", {k: list(v.shape) for k, v in feeds.items()}, ["scores"], feeds)
    np.savez_compressed(os.path.join(OUT, "cv2dnn_g4_lightglue.npz"), scores=o["scores"][0].astype(np.float32),
                        meta=json.dumps(dict(meta, inputs="synth.keypoint_set(160,752,480,7) / (144,...,8,perturb_of), normalize scale 0.5")))
    print("G4", o["scores"].shape)

    # ---- G5: SuperGlue indoor: everything up to the coupling matrix with dustbins ('2498'), then the 100 Sinkhorn iterations
    n0, n1 = host.normalize_keypoints(f0, 752, 480, 0.7), host.normalize_keypoints(f1, 752, 480, 0.7)
    feeds = {"keypoints_0": n0[1:3].T.copy()[None], "scores_0": n0[0][None].copy(), "descriptors_0": n0[3:][None].copy(),
             "keypoints_1": n1[1:3].T.copy()[None], "scores_1": n1[0][None].copy(), "descriptors_1": n1[3:][None].copy()}
    o = run("superglue_indoor_sim_int32.onnx", {k: list(v.shape) for k, v in feeds.items()}, ["2435"], feeds)
    sim = o["2435"].astype(np.float32).reshape(f0.shape[1], f1.shape[1])          # einsum(mdesc0, mdesc1) / sqrt(256), before the dustbins
    # The dustbin Concat behind this tensor is built from Shape / ConstantOfShape / Expand nodes that cv2.dnn cannot import: it is restated here
    # (three Expands of the scalar initialiser 'bin_score' + two Concats = nodes '2466'..'2498'; log_mu / log_nu = '2545' / '2559', norm = '2501')
    # and pinned against tools/onnx_interp.py (tests/test_oracle_golden.py).  The 100 Sinkhorn iterations themselves (200 ReduceLogSumExp + Sub /
    # Unsqueeze / Add nodes, '2563'..'scores') ARE executed by cv2.dnn, cut at those tensors with the nodes copied byte for byte.
    m, n = sim.shape
    bin_score = np.float32(R.load(REF + "superglue_indoor_sim_int32.onnx").init["bin_score"].reshape(()))
    z = np.full((m + 1, n + 1), bin_score, np.float32)
    z[:m, :n] = sim
    norm = -np.log(np.float32(m + n))
    log_mu = np.concatenate([np.full(m, norm, np.float32), [np.log(np.float32(n)) + norm]]).astype(np.float32)
    log_nu = np.concatenate([np.full(n, norm, np.float32), [np.log(np.float32(m)) + norm]]).astype(np.float32)
    feeds2 = {"2498": z[None], "2545": log_mu[None], "2559": log_nu[None], "2562": np.zeros((1, 1, n + 1), np.float32), "2501": np.array([norm], np.float32)}
    o2 = run("superglue_indoor_sim_int32.onnx", {k: list(v.shape) for k, v in feeds2.items()}, ["scores"], feeds2)
    scores = o2["scores"].astype(np.float32).reshape(m + 1, n + 1)
    np.savez_compressed(os.path.join(OUT, "cv2dnn_g5_superglue_indoor.npz"), sim=sim, couplings=z, scores=scores,
                        meta=json.dumps(dict(meta, inputs="same keypoint sets, normalize scale 0.7",
                                             sinkhorn="nodes '2563'..'scores' (100 iterations) executed by cv2.dnn on the couplings built from cv2.dnn's own similarity matrix")))
    print("G5", sim.shape)


if __name__ == "__main__":
    main()
