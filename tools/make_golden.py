"""Freeze golden vectors for the oracle from the reference's own ONNX graphs (run in the authoring container only).

    python -m tools.make_golden

Each fixture is produced by tools/onnx_interp.py executing /root/reference/output/*.onnx node by node in fp32 --
the reference's arithmetic specification (its TensorRT engines are built from exactly these files).  Weights are the
fp16-representable values shipped in weights/*.afw (what the reference's kFP16 engines hold); the drift against the
original fp32 initialisers is recorded in each file's `meta`.  Inputs are seeded synthetic data (oracle/synth.py).
Host-side stages (wireframe_matcher etc.) between graphs use oracle/host.py, as the reference uses its C++.
"""
import hashlib
import json
import os

import numpy as np
import torch

from oracle import host, synth
from . import onnx_interp as I
from . import onnx_reader as R

REF = "/root/reference/output/"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load(path, fp16_weights=True):
    g = R.load(REF + path)
    if fp16_weights:
        for k in list(g.init):
            v = g.init[k]
            if v.dtype == np.float32 and v.ndim >= 2 and v.size > 16 and "onnx::Mul" not in k:
                g.init[k] = v.astype(np.float16).astype(np.float32)
    return g


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    l, r, disp = synth.stereo_pair(752, 480, 0xA175)
    x = host.process_image(l)
    meta = {"image": "oracle.synth.stereo_pair(752, 480, 0xA175)[0]", "cv2": "resize restated, bit-exact vs cv2 4.13", "torch": torch.__version__}

    # ---- G1 SuperPoint
    o16 = I.Interp(load("superpoint_v1_sim_int32.onnx")).run({"input": x})
    o32 = I.Interp(load("superpoint_v1_sim_int32.onnx", False)).run({"input": x})
    sc, de = o16["scores"][0].numpy(), o16["descriptors"][0].numpy()
    nz = np.nonzero(sc.reshape(-1))[0].astype(np.int32)
    m = dict(meta, drift_vs_fp32_weights=dict(scores=float(np.abs(sc - o32["scores"][0].numpy()).max()),
                                               descriptors=float(np.abs(de - o32["descriptors"][0].numpy()).max())))
    np.savez_compressed(os.path.join(OUT, "g1_superpoint.npz"), nz_idx=nz, nz_val=sc.reshape(-1)[nz], desc_sha=sha(de),
                        desc_cells=de[:, ::8, ::8].copy(), meta=json.dumps(m))
    print("G1", len(nz), m["drift_vs_fp32_weights"])

    # ---- G2 PLNet s0 + host unique + G3
    g2 = I.Interp(load("plnet_s0.onnx")).run({"input": x})
    g2f = I.Interp(load("plnet_s0.onnx", False)).run({"input": x})
    keep_idx, inverse, pairs = host.wireframe_matcher(g2["iskeep"].numpy(), g2["idx_junc_to_end_min"].numpy(), g2["idx_junc_to_end_max"].numpy())
    feeds = dict(juncs_pred=g2["juncs_pred"].numpy(), lines_pred=g2["lines_pred"].numpy(), idx_lines_for_junctions=pairs.astype(np.float32),
                 inverse=inverse.astype(np.float32).reshape(-1, 1), iskeep_index=keep_idx.astype(np.float32).reshape(-1, 1),
                 loi_features=g2["loi_features"].numpy(), loi_features_thin=g2["loi_features_thin"].numpy(), loi_features_aux=g2["loi_features_aux"].numpy())
    g3 = I.Interp(load("plnet_s1.onnx")).run(feeds)
    m = dict(meta, drift_vs_fp32_weights=dict(
        lines_pred=float((g2["lines_pred"] - g2f["lines_pred"]).abs().max()), juncs_pred=float((g2["juncs_pred"] - g2f["juncs_pred"]).abs().max()),
        iskeep_flips=int((g2["iskeep"] != g2f["iskeep"]).sum()), loi=float((g2["loi_features"] - g2f["loi_features"]).abs().max())))
    np.savez_compressed(os.path.join(OUT, "g2_plnet_s0.npz"), juncs_pred=g2["juncs_pred"].numpy(), lines_pred_sub=g2["lines_pred"].numpy()[::97],
                        lines_pred_sha=sha(g2["lines_pred"].numpy()), iskeep_idx=keep_idx.astype(np.int32),
                        idx_min_kept=g2["idx_junc_to_end_min"].numpy()[keep_idx].astype(np.int16), idx_max_kept=g2["idx_junc_to_end_max"].numpy()[keep_idx].astype(np.int16),
                        loi_cells=g2["loi_features"].numpy()[0, :, ::16, ::16].copy(), thin=g2["loi_features_thin"].numpy()[0, :, ::4, ::4].copy(),
                        aux=g2["loi_features_aux"].numpy()[0, :, ::4, ::4].copy(), scores_nz=np.nonzero(g2["scores"].numpy().reshape(-1))[0].astype(np.int32),
                        meta=json.dumps(m))
    np.savez_compressed(os.path.join(OUT, "g3_plnet_s1.npz"), pairs=pairs.astype(np.int16), lines_adjusted=g3["lines_adjusted"].numpy(),
                        scores_line=g3["scores_line"].numpy(), meta=json.dumps(meta))
    print("G2/G3", len(keep_idx), len(pairs), m["drift_vs_fp32_weights"])
    # end-to-end PLNet outputs through the restated host code
    feats, lines, junc = host.plnet_process_output({k: v for k, v in g2.items()}, __import__("oracle.weights", fromlist=["x"]).load("plnet"),
                                                   host.PLNET_CFG_EUROC, 752, 480, True)
    np.savez_compressed(os.path.join(OUT, "plnet_e2e.npz"), feat_head=feats[:, :64], feat_xy=feats[:3], lines=lines, junc_xy=junc[:3], meta=json.dumps(meta))

    # ---- G4 / G5 matchers on planted correspondences
    f0 = synth.keypoint_set(160, 752, 480, 7)
    f1, perm = synth.keypoint_set(144, 752, 480, 8, perturb_of=f0)
    n0, n1 = host.normalize_keypoints(f0, 752, 480, 0.5), host.normalize_keypoints(f1, 752, 480, 0.5)
    feeds = {"keypoints_0": n0[1:3].T.copy()[None], "keypoints_1": n1[1:3].T.copy()[None], "descriptors_0": n0[3:].T.copy()[None], "descriptors_1": n1[3:].T.copy()[None]}
    s16 = I.Interp(load("superpoint_lightglue.onnx")).run(feeds)["scores"][0].numpy()
    s32 = I.Interp(load("superpoint_lightglue.onnx", False)).run(feeds)["scores"][0].numpy()
    idx, sc_m = host.filter_matches(s16)
    m = dict(meta, inputs="synth.keypoint_set(160,752,480,7) / (144,...,8,perturb_of)", drift_vs_fp32_weights=float(np.abs(np.exp(s16) - np.exp(s32)).max()))
    np.savez_compressed(os.path.join(OUT, "g4_lightglue.npz"), scores=s16.astype(np.float32), matches=idx, match_scores=sc_m, meta=json.dumps(m))
    n0, n1 = host.normalize_keypoints(f0, 752, 480, 0.7), host.normalize_keypoints(f1, 752, 480, 0.7)
    feeds = {"keypoints_0": n0[1:3].T.copy()[None], "scores_0": n0[0][None].copy(), "descriptors_0": n0[3:][None].copy(),
             "keypoints_1": n1[1:3].T.copy()[None], "scores_1": n1[0][None].copy(), "descriptors_1": n1[3:][None].copy()}
    for kind in ("indoor", "outdoor"):
        z = I.Interp(load("superglue_%s_sim_int32.onnx" % kind)).run(feeds)["scores"][0].numpy()
        i0, i1, m0, m1 = host.superglue_decode(z)
        np.savez_compressed(os.path.join(OUT, "g5_superglue_%s.npz" % kind), scores=z.astype(np.float32), indices0=i0, indices1=i1, mscores0=m0, mscores1=m1,
                            meta=json.dumps(meta))
    print("G4/G5 done", len(idx))


if __name__ == "__main__":
    main()
