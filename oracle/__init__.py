"""CPU oracle for the AirSLAM learned front-end (detect + match).

TEST INFRASTRUCTURE ONLY.  A CPU restatement (numpy / torch-CPU) of the
reference's algorithm for the hot path -- the five shipped ONNX graphs
 and the host code of src/plnet.cpp, src/super_point.cpp,
src/light_glue.cpp, src/super_glue.cpp, src/point_matcher.cc and
src/feature_detector.cc.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product path
(airslam_b200/) never does.

Pinning status: the reference ships no tests, golden vectors or fixtures for
this path and its arithmetic lives in TensorRT 8.6.1.6 (absent, closed source),
so the oracle is pinned against executions of the reference's own ONNX files:
(1) by an EXTERNAL runtime, OpenCV DNN 4.13, on static-shape sub-graphs cut byte
for byte out of the shipped files (tools/onnx_cut.py, tools/make_cv2dnn_golden.py
-> tests/golden/cv2dnn_*.npz, tests/test_oracle_cv2dnn.py): the dense part of
G1-G3 (incl. the in-graph NMS step by step, the HAFM line decode and the junction-heat NMS), the whole LightGlue graph, SuperGlue up to
the similarity matrix and its 100 Sinkhorn iterations;
(2) by the node-by-node interpreter tools/onnx_interp.py for what cv2.dnn cannot
import (integer / logical tails, the dustbin concat of SuperGlue) -> tests/golden/g*.npz.
With respect to the TensorRT engines themselves: PARITY UNPINNED.
"""
