"""The per-call paths of the class surface run as CUDA graphs from their third call on (first eager, second captured, see run_graphed in
csrc/airfe_capi.cu).  A replayed graph must give the bytes of the eager call: same kernels, same order -- for other inputs and other
keypoint counts of the same size bucket too."""
import numpy as np
import pytest

import _parity as P

pytestmark = pytest.mark.gpu


def _same_det(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


def test_detect_graph_replay_equals_eager():
    from airslam_b200 import capi
    from oracle import synth
    l1, r1, _ = synth.stereo_pair(752, 480, 401)
    l2, r2, _ = synth.stereo_pair(752, 480, 402)
    a = capi.Context(max_batch=1, enable_superpoint=0, enable_lightglue=0)
    b = capi.Context(max_batch=1, enable_superpoint=0, enable_lightglue=0)
    try:
        p1, p2 = np.stack([l1, r1]), np.stack([l2, r2])
        e1 = a.detect_batch(capi.NET_PLNET, p1, lines=True, junctions=True)     # eager
        c2 = a.detect_batch(capi.NET_PLNET, p2, lines=True, junctions=True)     # captured + launched
        g1 = a.detect_batch(capi.NET_PLNET, p1, lines=True, junctions=True)     # replay
        g2 = a.detect_batch(capi.NET_PLNET, p2, lines=True, junctions=True)     # replay
        f2 = b.detect_batch(capi.NET_PLNET, p2, lines=True, junctions=True)     # eager on a fresh context
        P.exact("graph: detect replay == eager (same frames)", all(_same_det(x, y) for x, y in zip(e1, g1)))
        P.exact("graph: detect capture / replay == eager of a fresh context (other frames)", all(_same_det(x, y) for x, y in zip(c2, f2)) and all(_same_det(x, y) for x, y in zip(g2, f2)))
        assert e1[0][0].shape[1] == 400 and e1[0][2].shape[1] > 10
    finally:
        a.close()
        b.close()


def test_match_graph_replay_equals_eager_across_counts_of_a_bucket():
    from airslam_b200 import capi
    from oracle import synth
    f0 = synth.keypoint_set(400, 752, 480, 11)
    f1, _ = synth.keypoint_set(400, 752, 480, 12, perturb_of=f0)
    g0 = synth.keypoint_set(390, 752, 480, 13)
    g1, _ = synth.keypoint_set(386, 752, 480, 14, perturb_of=g0)
    a = capi.Context(max_batch=1, enable_superpoint=0, enable_plnet=0)
    b = capi.Context(max_batch=1, enable_superpoint=0, enable_plnet=0)
    try:
        e = a.match_batch(capi.MATCHER_LIGHTGLUE, [f0], [f1])[0]      # eager
        a.match_batch(capi.MATCHER_LIGHTGLUE, [f0], [f1])             # capture
        r = a.match_batch(capi.MATCHER_LIGHTGLUE, [f0], [f1])[0]      # replay, same input
        r2 = a.match_batch(capi.MATCHER_LIGHTGLUE, [g0], [g1])[0]     # replay, other counts of the same bucket (stale rows beyond n in the staging)
        f = b.match_batch(capi.MATCHER_LIGHTGLUE, [g0], [g1])[0]      # eager, fresh context
        P.exact("graph: match replay == eager", np.array_equal(e[0], r[0]) and np.array_equal(e[1], r[1]) and len(e[0]) > 300)
        P.exact("graph: match replay with other keypoint counts of the bucket == eager", np.array_equal(r2[0], f[0]) and np.array_equal(r2[1], f[1]) and len(f[0]) > 300)
    finally:
        a.close()
        b.close()
