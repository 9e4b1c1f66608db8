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
so the oracle is pinned against the next best thing: a node-by-node fp32
execution of the reference's own ONNX files (tools/onnx_interp.py, run in the
authoring container) frozen under tests/golden/.  With respect to the
TensorRT engines themselves: PARITY UNPINNED.
"""
