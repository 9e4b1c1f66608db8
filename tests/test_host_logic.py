"""CPU: restated host semantics of the reference (tie-breaks, inclusive borders, integer divisions, empty inputs)."""
import numpy as np
import pytest

from oracle import host, synth


def test_cv_resize_bit_exact_against_cv2():
    cv2 = pytest.importorskip("cv2")
    for (w, h) in ((752, 480), (640, 480), (1280, 720), (376, 240), (512, 512)):
        img = synth.stereo_pair(w, h, 3)[0]
        assert np.array_equal(host.cv_resize_u8(img), cv2.resize(img, (512, 512))), (w, h)
    noise = np.random.RandomState(0).randint(0, 256, (480, 752)).astype(np.uint8)
    assert np.array_equal(host.cv_resize_u8(noise), cv2.resize(noise, (512, 512)))


def test_process_image_is_double_division_then_float():
    img = np.arange(512 * 512, dtype=np.uint32).reshape(512, 512).astype(np.uint8)
    x = host.process_image(img)
    assert x.dtype == np.float32 and x.shape == (1, 1, 512, 512)
    assert np.array_equal(x[0, 0], (img.astype(np.float64) / 255.0).astype(np.float32))


def test_detect_point_border_is_inclusive_and_ties_are_raster_ordered():
    heat = np.zeros((512, 512), dtype=np.float32)
    heat[4, 4] = 0.5       # x == border: kept
    heat[3, 100] = 0.9     # y < border: dropped
    heat[508, 508] = 0.5   # x == W - border: kept (reference uses '>' W-border, src/plnet.cpp:332)
    heat[509, 10] = 0.9    # y > H - border: dropped
    heat[100, 100] = 0.5
    pts = host.detect_point(heat, 0.004, 4, 400)
    assert pts.shape == (3, 3)
    assert [tuple(p) for p in pts[1:].T] == [(4.0, 4.0), (100.0, 100.0), (508.0, 508.0)]    # raster order when <= top_k
    pts = host.detect_point(heat, 0.004, 4, 2)
    assert [tuple(p) for p in pts[1:].T] == [(4.0, 4.0), (100.0, 100.0)]                   # equal scores: lower raster index first
    assert host.detect_point(np.zeros((512, 512), np.float32), 0.004, 4, 400).shape == (3, 0)


def test_extract_descriptors_unit_norm_and_zero_column():
    rs = np.random.RandomState(1)
    desc = rs.normal(size=(256, 64, 64)).astype(np.float32)
    desc[:, :3, :3] = 0
    pts = np.array([[0.1, 0.2, 0.3], [250.0, 4.0, 508.0], [300.0, 4.0, 4.0]], dtype=np.float32)
    d = host.extract_descriptors(desc, pts)
    assert d.shape == (256, 3)
    assert abs(np.linalg.norm(d[:, 0]) - 1) < 1e-5 and abs(np.linalg.norm(d[:, 2]) - 1) < 1e-5
    assert np.all(d[:, 1] == 0)      # all-zero neighbourhood stays zero (Eigen normalize() guard), no NaN


def test_wireframe_matcher_first_seen_order_and_swap():
    keep = np.array([0, 1, 1, 0, 1, 1, 1], dtype=np.float32)
    imin = np.array([0, 5, 2, 0, 5, 1, 2], dtype=np.float32)
    imax = np.array([0, 9, 7, 0, 9, 3, 7], dtype=np.float32)
    ki, inv, pairs = host.wireframe_matcher(keep, imin, imax)
    assert ki.tolist() == [1, 2, 4, 5, 6]
    assert inv.tolist() == [0, 1, 0, 2, 1]
    assert pairs.tolist() == [[9, 5], [7, 2], [3, 1]]      # (max, min)
    ki, inv, pairs = host.wireframe_matcher(np.zeros(4, np.float32), np.zeros(4, np.float32), np.zeros(4, np.float32))
    assert len(ki) == 0 and pairs.shape == (0, 2)


def test_normalize_keypoints_integer_half_width():
    f = np.zeros((259, 2), dtype=np.float32)
    f[1] = [376.0, 0.0]
    f[2] = [240.0, 10.0]
    n = host.normalize_keypoints(f, 753, 481, 0.5)      # 753 / 2 == 376 (integer division, src/point_matcher.cc:45)
    assert n[1, 0] == 0.0 and n[2, 0] == 0.0
    l_inv = np.float32(1.0 / 753 * np.float64(np.float32(0.5)))
    assert n[1, 1] == np.float32(-376.0) * l_inv


def test_filter_matches_first_max_and_threshold():
    s = np.full((3, 4), -10.0, dtype=np.float32)
    s[0, 1] = s[0, 2] = -0.1          # row tie: first column wins
    s[1, 1] = -0.05                   # column 1's max is row 1 -> row 0 is not mutual
    s[2, 3] = np.log(0.1)             # exp == threshold exactly: rejected ('>' 0.1)
    idx, sc = host.filter_matches(s)
    assert idx.tolist() == [[1, 1]]
    assert abs(sc[0] - np.exp(np.float32(-0.05))) < 1e-7


def test_superglue_decode_ignores_dustbins_and_thresholds():
    z = np.full((3, 3), -20.0, dtype=np.float32)
    z[0, 1] = np.log(0.9)
    z[1, 0] = np.log(0.15)            # mutual but exp <= 0.2 -> invalid
    z[2, :] = 5.0                     # dustbin row / col are ignored
    z[:, 2] = 5.0
    i0, i1, m0, m1 = host.superglue_decode(z)
    assert i0.tolist() == [1, -1] and i1.tolist() == [-1, 0]
    assert abs(m0[0] - 0.9) < 1e-6 and abs(m0[1] - 0.15) < 1e-6 and abs(m1[1] - 0.9) < 1e-6


def test_matching_points_empty_side():
    f = synth.keypoint_set(10, 752, 480, 1)
    assert host.matching_points(f, np.zeros((259, 0), np.float32), {}, 0, 752, 480) == []


def _assoc_full(lines, juncs):
    """The graph's association (G2, plnet.cpp:453-462 outputs): first-minimum ArgMin over all junctions, fp32 arithmetic."""
    def nearest(ex, ey):
        dx = (ex[:, None] - juncs[None, :, 0]).astype(np.float32)
        dy = (ey[:, None] - juncs[None, :, 1]).astype(np.float32)
        d = (dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)
        return d.min(1), d.argmin(1)
    m1, i1 = nearest(lines[:, 0], lines[:, 1])
    m2, i2 = nearest(lines[:, 2], lines[:, 3])
    lo, hi = np.minimum(i1, i2), np.maximum(i1, i2)
    return lo, hi, (lo < hi) & (m1 < 10.0) & (m2 < 10.0)


def _assoc_grid(lines, juncs, cell_cap=8):
    """What assoc_kernel (csrc/line_kernels.cu) does: 32x32 grid of 4x4-px cells, 3x3 neighbourhood around each endpoint,
    candidates ordered by (distance, index); returns None when a cell overflows (the kernel then runs the full scan)."""
    cells = {}
    for q, (x, y) in enumerate(juncs):
        c = (min(31, max(0, int(np.float32(y) * np.float32(0.25)))), min(31, max(0, int(np.float32(x) * np.float32(0.25)))))
        cells.setdefault(c, []).append(q)
    if max(len(v) for v in cells.values()) > cell_cap:
        return None
    n = len(lines)
    lo, hi, keep = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, bool)

    def nearest(ex, ey):
        best, bi = np.float32(np.inf), 0
        cx0, cx1 = max(0, int(np.floor((ex - 3.17) * 0.25))), min(31, int(np.floor((ex + 3.17) * 0.25)))
        cy0, cy1 = max(0, int(np.floor((ey - 3.17) * 0.25))), min(31, int(np.floor((ey + 3.17) * 0.25)))
        for cy in range(cy0, cy1 + 1):
            for cx in range(cx0, cx1 + 1):
                for q in cells.get((cy, cx), ()):
                    dx, dy = np.float32(ex - juncs[q, 0]), np.float32(ey - juncs[q, 1])
                    d = np.float32(dx * dx) + np.float32(dy * dy)
                    if d < best or (d == best and q < bi):
                        best, bi = d, q
        return best, bi
    for k in range(n):
        m1, i1 = nearest(lines[k, 0], lines[k, 1])
        m2, i2 = nearest(lines[k, 2], lines[k, 3])
        if not (m1 < 10.0 and m2 < 10.0):
            continue
        lo[k], hi[k] = min(i1, i2), max(i1, i2)
        keep[k] = lo[k] < hi[k]
    return lo, hi, keep


def test_association_grid_pruning_equals_full_scan_on_kept_rows():
    """The device prunes the 300-junction scan to the sqrt(10)-px neighbourhood: the keep mask must be identical for every proposal and
    (imin, imax) identical wherever keep is set -- the only rows wireframe_matcher reads (plnet.cpp:272-307)."""
    rng = np.random.RandomState(5)
    for trial in range(3):
        # junctions: distinct integer cells + sub-pixel offsets, like TopK over a 3x3-NMS'd map (plus exact-tie bait: duplicated coordinates)
        cellsel = rng.choice(128 * 128, 300, replace=False)
        juncs = np.stack([(cellsel % 128) + 0.5 + rng.uniform(-0.49, 0.49, 300), (cellsel // 128) + 0.5 + rng.uniform(-0.49, 0.49, 300)], 1).astype(np.float32)
        juncs[7] = juncs[3]                      # an exact tie: the first index must win
        # proposals: half of them start / end near junctions so that many are kept
        n = 1500
        a, b = rng.randint(0, 300, n), rng.randint(0, 300, n)
        lines = np.concatenate([juncs[a] + rng.normal(0, 1.5, (n, 2)), juncs[b] + rng.normal(0, 1.5, (n, 2))], 1)
        lines[n // 2:] = rng.uniform(0, 127, (n - n // 2, 4))
        lines = np.clip(lines, 0, 127).astype(np.float32)
        lo_f, hi_f, keep_f = _assoc_full(lines, juncs)
        got = _assoc_grid(lines, juncs)
        assert got is not None
        lo_g, hi_g, keep_g = got
        assert np.array_equal(keep_f, keep_g)
        assert keep_f.sum() > 100
        assert np.array_equal(lo_f[keep_f], lo_g[keep_f]) and np.array_equal(hi_f[keep_f], hi_g[keep_f])
    # TopK padding with non-peak cells produces runs of adjacent junctions: the grid must refuse (the kernel falls back to the full scan)
    dense = np.stack([np.arange(300) % 128 + 0.5, np.arange(300) // 128 + 0.5], 1).astype(np.float32)
    assert _assoc_grid(lines, dense) is None


def test_cv_remap_restatement_is_bit_exact_against_cv2():
    """oracle.host.cv_remap_u8 == cv2.remap(INTER_LINEAR) on 8-bit gray with CV_32F maps (src/camera.cc:161-182), inside the image and
    across its borders (BORDER_CONSTANT 0); the synthetic radial-tangential maps agree with cv2.initUndistortRectifyMap."""
    cv2 = pytest.importorskip("cv2")
    from oracle import host, synth
    img, _, _ = synth.stereo_pair(752, 480, 77)
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]])
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0])
    newK = K.copy()
    newK[0, 2], newK[1, 2] = 362.0, 250.0
    m1, m2 = cv2.initUndistortRectifyMap(K, D, np.eye(3), newK, (752, 480), cv2.CV_32F)
    assert np.array_equal(host.cv_remap_u8(img, m1, m2), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))
    for sx, ox, sy, oy in ((1.3, -100.0, 1.2, -40.0), (0.5, 300.0, 0.5, 200.0), (1.0, 0.499, 1.0, -0.501)):      # far outside / magnified / half-pixel ties
        a, b = (m1 * sx + ox).astype(np.float32), (m2 * sy + oy).astype(np.float32)
        assert np.array_equal(host.cv_remap_u8(img, a, b), cv2.remap(img, a, b, cv2.INTER_LINEAR))
    mx, my = host.radtan_rectify_maps(752, 480, 458.654, 457.296, 367.215, 248.375, *D[:4], new_cx=362.0, new_cy=250.0)
    assert np.abs(mx - m1).max() < 2e-3 and np.abs(my - m2).max() < 2e-3


def test_line_association_restatement_small_case():
    """oracle.host.assign_points_to_lines / match_lines (src/line_processor.cc:68-187) on a case small enough to check by hand."""
    from oracle import host
    lines0 = np.array([[10.0, 10.0, 110.0, 10.0], [50.0, 0.0, 50.0, 100.0]])           # horizontal, vertical
    lines1 = np.array([[5.0, 10.0, 105.0, 10.0], [45.0, 0.0, 45.0, 100.0]])            # the same, shifted 5 px left
    f0 = np.zeros((259, 6), np.float32)
    f0[1], f0[2] = [20, 60, 100, 50, 50, 300], [11, 9, 10, 40, 80, 300]                 # 3 on line 0, 2 on line 1 (+ (50,10)?) , 1 nowhere
    f1 = f0.copy()
    f1[1] -= 5
    rel0, rel1 = host.assign_points_to_lines(lines0, f0), host.assign_points_to_lines(lines1, f1)
    assert [sorted(r) for r in rel0] == [[0, 1, 2], [3, 4]] and [sorted(r) for r in rel1] == [[0, 1, 2], [3, 4]]
    assert rel0[0][0] == 1.0 and rel0[0][1] == 1.0 and rel0[0][2] == 0.0 and rel0[1][3] == 0.0
    # point (113, 10): beyond the end point but within 3 px -> side2 <= 9 keeps it; (118, 10): outside the +3 box
    g = np.zeros((259, 2), np.float32)
    g[1], g[2] = [113, 118], [10, 10]
    assert [sorted(r) for r in host.assign_points_to_lines(lines0[:1], g)] == [[0]]
    matches = [(j, j) for j in range(6)]
    kept = host.filter_stereo_matches(f0, f1, matches, 2.0, 50.0, 2.0)
    assert kept == matches                                                                # disparity 5 px, dy 0
    assert host.filter_stereo_matches(f0, f1, matches, 6.0, 50.0, 2.0) == []
    assert host.match_lines(rel0, rel1, kept, 6, 6) == [0, 1]                             # 3 votes / 3 points and 2 votes / 2 points
    assert host.match_lines(rel0, rel1, kept[:1] + kept[3:], 6, 6) == [-1, 1]             # one vote is not enough (col_max_val < 2)
