"""Oracle restatement of the reference's HOST code on the hot path (numpy).  TEST INFRASTRUCTURE.

Every function cites the reference lines it follows (paths relative to /root/reference).
Feature matrices are numpy float32 arrays of shape [259, N]: row 0 score, rows 1-2 x,y, rows 3..258
descriptor -- the same column-per-keypoint convention as Eigen::Matrix<float,259,Dynamic>.
"""
import numpy as np

from . import nets

RESIZED = 512      # src/plnet.cpp:17-18, src/super_point.cpp:12-13


# ---- cv::resize(CV_8U, INTER_LINEAR) -- integer formula of SURVEY.md App. B9 ---------------------
def cv_resize_u8(img, dst_w=RESIZED, dst_h=RESIZED):
    """Bit-exact restatement of cv::resize on 8-bit gray (src/plnet.cpp:258, src/super_point.cpp:116).
    Checked against cv2 4.13 in tests/test_oracle_host.py."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    src_h, src_w = img.shape
    sx_scale = np.float64(src_w) / dst_w
    sy_scale = np.float64(src_h) / dst_h

    def coeffs(n_dst, n_src, scale, reset_border):
        d = np.arange(n_dst, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int32)
        f = (f - s.astype(np.float32)).astype(np.float32)
        if reset_border:
            lo = s < 0
            f[lo] = 0
            s[lo] = 0
            hi = s >= n_src - 1
            f[hi] = 0
            s[hi] = n_src - 1
        a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
        a1 = np.rint(f * np.float32(2048)).astype(np.int32)
        return s, a0, a1

    sx, a0, a1 = coeffs(dst_w, src_w, sx_scale, True)
    sx1 = np.minimum(sx + 1, src_w - 1)
    s32 = img.astype(np.int32)
    hbuf = s32[:, sx] * a0[None, :] + s32[:, sx1] * a1[None, :]          # [src_h, dst_w] int32
    sy, b0, b1 = coeffs(dst_h, src_h, sy_scale, False)
    r0 = np.clip(sy, 0, src_h - 1)
    r1 = np.clip(sy + 1, 0, src_h - 1)
    h0 = hbuf[r0] >> 4
    h1 = hbuf[r1] >> 4
    out = (((b0[:, None] * h0) >> 16) + ((b1[:, None] * h1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def cv_remap_u8(img, map_x, map_y):
    """Camera::UndistortImage, src/camera.cc:161-182: cv::remap(img, rect, map1, map2, cv::INTER_LINEAR) with CV_32F maps (camera.cc:63-66),
    8-bit gray, default BORDER_CONSTANT(0).  Restated from OpenCV's remap: the float maps are quantised to 1/32 pixel (cvRound(32 x), round
    half to even), the four bilinear weights are the 15-bit table entries (32-fx)(32-fy)*32 ... (exact, they sum to 2^15), and the result is
    (sum + 2^14) >> 15.  Bit-exact against cv2 4.13 (tests/test_host_logic.py), including samples that leave the image."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    sx = np.rint(np.asarray(map_x, dtype=np.float32) * np.float32(32)).astype(np.int64)
    sy = np.rint(np.asarray(map_y, dtype=np.float32) * np.float32(32)).astype(np.int64)
    fx, fy = sx & 31, sy & 31
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)

    def px(y, x):
        ok = (x >= 0) & (x < w) & (y >= 0) & (y < h)
        return np.where(ok, img[np.clip(y, 0, h - 1), np.clip(x, 0, w - 1)].astype(np.int64), 0)
    v = (px(iy, ix) * ((32 - fx) * (32 - fy) * 32) + px(iy, ix + 1) * (fx * (32 - fy) * 32) + px(iy + 1, ix) * ((32 - fx) * fy * 32) +
         px(iy + 1, ix + 1) * (fx * fy * 32) + (1 << 14)) >> 15
    return np.clip(v, 0, 255).astype(np.uint8)


def radtan_rectify_maps(w, h, fx, fy, cx, cy, k1, k2, p1, p2, k3=0.0, new_cx=None, new_cy=None, new_f_scale=1.0):
    """Synthetic stand-in for cv::initUndistortRectifyMap(K, D, I, K', size, CV_32F) (src/camera.cc:63-66) with R = identity: for every
    rectified pixel the position in the raw (radial-tangential distorted) image.  Test / bench input generator, not on the hot path."""
    new_cx = cx if new_cx is None else new_cx
    new_cy = cy if new_cy is None else new_cy
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    x, y = (u - new_cx) / (fx * new_f_scale), (v - new_cy) / (fy * new_f_scale)     # new_f_scale < 1: the rectified view is wider than the sensor
    r2 = x * x + y * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * kr + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return (fx * xd + cx).astype(np.float32), (fy * yd + cy).astype(np.float32)


def process_image(img):
    """src/plnet.cpp:246-270 == src/super_point.cpp:111-116,146-165: resize to 512x512, float(u8)/255.0 (double) -> f32."""
    r = cv_resize_u8(img)
    return (r.astype(np.float64) / 255.0).astype(np.float32).reshape(1, 1, RESIZED, RESIZED)


# ---- keypoints ----------------------------------------------------------------------------------
def detect_point(heat, threshold, border, top_k):
    """src/plnet.cpp:309-355 == src/super_point.cpp:174-217.  heat [H,W] f32 -> [3,N] (score,x,y).
    Tie order of the reference's std::sort is unspecified; this build defines it as
    score descending, raster index ascending (SURVEY.md §8a hazards)."""
    h, w = heat.shape
    flat = heat.reshape(-1)
    idx = np.nonzero(flat >= np.float32(threshold))[0]
    y = idx // w
    x = idx - y * w
    ok = ~((x < border) | (x > w - border) | (y < border) | (y > h - border))   # inclusive upper bound
    idx, x, y = idx[ok], x[ok], y[ok]
    sc = flat[idx]
    if len(idx) > top_k:
        order = np.argsort(-sc, kind="stable")[:top_k]
        sc, x, y = sc[order], x[order], y[order]
    return np.stack([sc, x.astype(np.float32), y.astype(np.float32)]).astype(np.float32)


def extract_descriptors(desc, pts, h=64, w=64, s=8):
    """src/plnet.cpp:369-417 == src/super_point.cpp:224-272.  desc [256,h,w] f32; pts [3,N] -> [256,N] unit columns."""
    f32 = np.float32
    den = w * s - s // 2 - 0.5                      # s/2 is integer division in the reference
    sx = f32(2.0) / f32(den)
    bx = f32((1 - s) / den - 1)
    den_y = h * s - s // 2 - 0.5
    sy = f32(2.0) / f32(den_y)
    by = f32((1 - s) / den_y - 1)
    xs = ((pts[1].astype(f32) * sx + bx) + f32(1)) * f32(0.5)
    ys = ((pts[2].astype(f32) * sy + by) + f32(1)) * f32(0.5)
    ix = (xs * f32(w - 1)).astype(f32)
    iy = (ys * f32(h - 1)).astype(f32)

    def clip(v, m):
        return np.minimum(np.maximum(v, 0), m - 1)
    ix_nw = clip(np.floor(ix).astype(np.int64), w)
    iy_nw = clip(np.floor(iy).astype(np.int64), h)
    ix_ne, iy_ne = clip(ix_nw + 1, w), clip(iy_nw, h)
    ix_sw, iy_sw = clip(ix_nw, w), clip(iy_nw + 1, h)
    ix_se, iy_se = clip(ix_nw + 1, w), clip(iy_nw + 1, h)
    nw = (ix_se.astype(f32) - ix) * (iy_se.astype(f32) - iy)
    ne = (ix - ix_sw.astype(f32)) * (iy_sw.astype(f32) - iy)
    sw = (ix_ne.astype(f32) - ix) * (iy - iy_ne.astype(f32))
    se = (ix - ix_nw.astype(f32)) * (iy - iy_nw.astype(f32))
    d = desc.astype(f32)
    out = (d[:, iy_nw, ix_nw] * nw + d[:, iy_ne, ix_ne] * ne + d[:, iy_sw, ix_sw] * sw + d[:, iy_se, ix_se] * se).astype(f32)
    n2 = (out.astype(f32) ** 2).sum(axis=0, dtype=f32)
    nz = n2 > 0                                      # Eigen normalize(): leaves all-zero columns untouched
    out[:, nz] = out[:, nz] / np.sqrt(n2[nz])
    return out


def keypoints_decoder(scores, desc, threshold, border, top_k):
    """src/plnet.cpp:419-423.  Returns [259,N] in 512-space (no rescale here)."""
    pts = detect_point(scores, threshold, border, top_k)
    d = extract_descriptors(desc, pts)
    return np.concatenate([pts, d], axis=0).astype(np.float32)


# ---- PLNet line path ----------------------------------------------------------------------------
def wireframe_matcher(iskeep, idx_min, idx_max):
    """src/plnet.cpp:272-307.  Returns (is_keep_index, inverse, idx_lines_for_junctions_unique [(max,min)])."""
    keep_idx = np.nonzero(iskeep > 0)[0]
    seen = {}
    inverse = np.empty(len(keep_idx), dtype=np.int64)
    pairs = []
    for k, i in enumerate(keep_idx):
        key = (int(idx_min[i]), int(idx_max[i]))
        uid = seen.get(key)
        if uid is None:
            uid = len(pairs)
            seen[key] = uid
            pairs.append((key[1], key[0]))            # (max, min): the swap at plnet.cpp:301
        inverse[k] = uid
    return keep_idx.astype(np.int64), inverse, np.array(pairs, dtype=np.int64).reshape(-1, 2)


def plnet_process_output(out0, weights, cfg, in_w, in_h, junction_detection, emul=False, keep=None):
    """src/plnet.cpp:450-585.  out0: dict of the 10 stage-0 outputs (numpy/torch).  Returns (features [259,N],
    lines [L,4] float64, junctions [259,J])."""
    g = {k: (v.numpy() if hasattr(v, "numpy") else np.asarray(v)) for k, v in out0.items()}
    keep_idx, inverse, pairs = wireframe_matcher(g["iskeep"], g["idx_junc_to_end_min"], g["idx_junc_to_end_max"])
    w_scale = np.float32(in_w) / np.float32(RESIZED)
    h_scale = np.float32(in_h) / np.float32(RESIZED)
    lines = []
    junction_map = np.zeros((RESIZED, RESIZED), dtype=bool)
    if len(pairs) > 0:
        adj, score = nets.plnet_s1_forward(g["juncs_pred"], g["lines_pred"], pairs.astype(np.float32),
                                           inverse.astype(np.float32), keep_idx.astype(np.float32),
                                           g["loi_features"], g["loi_features_thin"], g["loi_features_aux"],
                                           weights, emul=emul, keep=keep)
        adj, score = adj.numpy(), score.numpy()
        if keep is not None:
            keep.update(lines_adjusted=adj, scores_line=score, pairs=pairs, inverse=inverse, keep_idx=keep_idx)
        border = max(cfg["remove_borders"], 0)
        len2_thr = np.float32(cfg["line_length_threshold"]) * np.float32(cfg["line_length_threshold"])
        for i in range(len(pairs)):
            if score[i] < 0.5:
                continue
            x1, y1, x2, y2 = (np.float32(adj[i, k]) * np.float32(4) for k in range(4))
            xi1, yi1, xi2, yi2 = int(x1 + np.float32(0.1)), int(y1 + np.float32(0.1)), int(x2 + np.float32(0.1)), int(y2 + np.float32(0.1))
            p1 = (xi1 > border) and (xi1 < RESIZED - border) and (yi1 > border) and (yi1 < RESIZED - border)
            p2 = (xi2 > border) and (xi2 < RESIZED - border) and (yi2 > border) and (yi2 < RESIZED - border)
            # the reference writes without bounds checks (plnet.cpp:540-541); coordinates are <= 511.5 so in range
            junction_map[min(yi1, RESIZED - 1), min(xi1, RESIZED - 1)] = p1
            junction_map[min(yi2, RESIZED - 1), min(xi2, RESIZED - 1)] = p2
            if score[i] < np.float32(cfg["line_threshold"]):
                continue
            l2 = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1)
            if l2 < len2_thr:
                continue
            lines.append((float(x1), float(y1), float(x2), float(y2)))
    scores = g["scores"].reshape(RESIZED, RESIZED)
    desc = g["descriptors"].reshape(256, 64, 64)
    feats = keypoints_decoder(scores, desc, cfg["keypoint_threshold"], cfg["remove_borders"], cfg["max_keypoints"])
    junctions = np.zeros((259, 0), dtype=np.float32)
    if junction_detection:
        border = max(cfg["remove_borders"], 0)
        sub = junction_map[border:RESIZED - border, border:RESIZED - border]
        yy, xx = np.nonzero(sub)                                # raster order (plnet.cpp:429-437)
        yy, xx = yy + border, xx + border
        pts = np.stack([scores[yy, xx], xx.astype(np.float32), yy.astype(np.float32)]).astype(np.float32)
        junctions = np.concatenate([pts, extract_descriptors(desc, pts)], axis=0).astype(np.float32)
        junctions[1] *= w_scale
        junctions[2] *= h_scale
    feats[1] *= w_scale
    feats[2] *= h_scale
    lines = np.array(lines, dtype=np.float64).reshape(-1, 4)
    lines[:, 0] *= w_scale
    lines[:, 1] *= h_scale
    lines[:, 2] *= w_scale
    lines[:, 3] *= h_scale
    return feats, lines, junctions


PLNET_CFG_EUROC = dict(max_keypoints=400, keypoint_threshold=0.004, remove_borders=4,
                       line_threshold=0.75, line_length_threshold=50.0)   # configs/visual_odometry/vo_euroc.yaml:2-7


def plnet_infer(img, weights, cfg, junction_detection=False, emul=False, keep=None):
    """PLNet::infer, src/plnet.cpp:221-244."""
    x = process_image(img)
    out0 = nets.plnet_s0_forward(x, weights, emul=emul, keep=keep)
    return plnet_process_output(out0, weights, cfg, img.shape[1], img.shape[0], junction_detection, emul=emul, keep=keep)


def superpoint_infer(img, weights, cfg, emul=False, prefix="sp.", keep=None):
    """SuperPoint::infer, src/super_point.cpp:103-144 + :274-282 (rescale inside keypoints_decoder)."""
    x = process_image(img)
    scores, desc = nets.superpoint_forward(x, weights, prefix=prefix, emul=emul, keep=keep)
    feats = keypoints_decoder(scores[0].numpy(), desc[0].numpy(), cfg["keypoint_threshold"], cfg["remove_borders"], cfg["max_keypoints"])
    feats[1] *= np.float32(img.shape[1]) / np.float32(RESIZED)
    feats[2] *= np.float32(img.shape[0]) / np.float32(RESIZED)
    return feats


# ---- matcher host code --------------------------------------------------------------------------
def normalize_keypoints(feats, width, height, scale):
    """PointMatcher::NormalizeKeypoints, src/point_matcher.cc:39-48 (width/2 integer division; multiplies by scale)."""
    out = feats.astype(np.float32).copy()
    l_inv = np.float32(1.0 / max(width, height) * np.float64(np.float32(scale)))
    out[1] = (feats[1] - np.float32(width // 2)) * l_inv
    out[2] = (feats[2] - np.float32(height // 2)) * l_inv
    return out


def filter_matches(scores, threshold=0.1):
    """src/light_glue.cpp:214-266: first-max row/col argmax, mutual check, exp(score) > threshold."""
    s = np.asarray(scores, dtype=np.float32)
    rmax = np.argmax(s, axis=1)             # numpy argmax returns the first maximum (strict '>' scan)
    cmax = np.argmax(s, axis=0)
    rows = np.arange(s.shape[0])
    mutual = cmax[rmax] == rows
    e = np.exp(s[rows, rmax]).astype(np.float32)
    ok = mutual & (e > np.float32(threshold))
    idx = np.stack([rows[ok], rmax[ok]], axis=1).astype(np.int32)
    return idx, e[ok]


def lightglue_infer(f0_258, f1_258, weights, emul=False, keep=None):
    """SuperPointLightGlue::infer, src/light_glue.cpp:120-170 (+process_input :172-212, process_output :268-281).
    f*_258: [258,N] (rows 0-1 keypoints, 2..257 descriptors)."""
    k0, k1 = f0_258[0:2].T.copy(), f1_258[0:2].T.copy()
    d0, d1 = f0_258[2:].T.copy(), f1_258[2:].T.copy()
    sc = nets.lightglue_forward(k0, k1, d0, d1, weights, emul=emul, keep=keep).numpy()
    return filter_matches(sc) + (sc,)


def superglue_decode(sc, threshold=0.2):
    """decode(), src/super_glue.cpp:339-367 with max_matrix / where helpers :248-337."""
    s = np.asarray(sc, dtype=np.float32)[:-1, :-1]
    m, n = s.shape
    max0, idx0 = s.max(axis=1), np.argmax(s, axis=1)
    idx1 = np.argmax(s, axis=0)
    mutual0 = np.arange(m) == idx1[idx0]
    mutual1 = np.arange(n) == idx0[idx1]
    ms0 = np.where(mutual0, np.exp(max0.astype(np.float32)), np.float32(0)).astype(np.float32)
    ms1 = np.where(mutual1, ms0[idx1], np.float32(0)).astype(np.float32)
    valid0 = mutual0 & (ms0 > np.float32(threshold))
    valid1 = mutual1 & valid0[idx1]
    return (np.where(valid0, idx0, -1).astype(np.int32), np.where(valid1, idx1, -1).astype(np.int32),
            ms0.astype(np.float64), ms1.astype(np.float64))   # VectorXd at the class surface (super_glue.cpp:462-469)


def superglue_infer(f0_259, f1_259, weights, emul=False, keep=None):
    """SuperGlue::infer, src/super_glue.cpp:137-197 (process_input :199-246)."""
    sc = nets.superglue_forward(f0_259[1:3].T.copy(), f0_259[0].copy(), f0_259[3:].copy(),
                                f1_259[1:3].T.copy(), f1_259[0].copy(), f1_259[3:].copy(), weights, emul=emul, keep=keep).numpy()
    return superglue_decode(sc) + (sc,)


def matching_points(f0, f1, weights, matcher, image_width, image_height, emul=False, keep=None):
    """PointMatcher::MatchingPoints, src/point_matcher.cc:50-108 with outlier_rejection=false.
    Returns list of (queryIdx, trainIdx, distance)."""
    if f0.shape[1] < 1 or f1.shape[1] < 1:
        return []
    scale = 0.7 if matcher else 0.5
    n0 = normalize_keypoints(f0, image_width, image_height, scale)
    n1 = normalize_keypoints(f1, image_width, image_height, scale)
    out = []
    if matcher == 0:
        idx, score, _ = lightglue_infer(n0[1:], n1[1:], weights, emul=emul, keep=keep)
        for (i, j), s in zip(idx, score):
            out.append((int(i), int(j), float(np.float32(1.0 - np.float64(s)))))
    else:
        i0, i1, ms0, ms1, _ = superglue_infer(n0, n1, weights, emul=emul, keep=keep)
        for i in range(len(i0)):
            if 0 <= i0[i] < len(i1) and i1[i0[i]] == i:
                out.append((i, int(i0[i]), float(np.float32(1.0 - (ms0[i] + ms1[i0[i]]) / 2.0))))
    return out


# ---- point <-> line association and stereo line matching --------------------
def assign_points_to_lines(lines, feats):
    """AssignPointsToLines, src/line_processor.cc:68-120.  lines [L,4] float64 (x1,y1,x2,y2), feats [259,N] float32.
    Returns a list (one entry per line) of dicts {point index: distance} -- std::map<int,double> holding double(float(distance)).
    std::pow(x, 2) is restated as x * x (exact square; glibc's pow returns the same value)."""
    lines = np.asarray(lines, dtype=np.float64).reshape(-1, 4)
    px = feats[1].astype(np.float64)
    py = feats[2].astype(np.float64)
    rel = []
    for (lx1, ly1, lx2, ly2) in lines:
        A, B = ly2 - ly1, lx1 - lx2
        C = lx2 * ly1 - lx1 * ly2
        D = np.sqrt(A * A + B * B)
        min_lx, max_lx = (lx1, lx2) if lx1 <= lx2 else (lx2, lx1)
        min_ly, max_ly = (ly1, ly2) if ly1 <= ly2 else (ly2, ly1)
        box = ~((px < min_lx - 3) | (px > max_lx + 3) | (py < min_ly - 3) | (py > max_ly + 3))
        with np.errstate(divide="ignore", invalid="ignore"):
            dist = (np.abs((A * px + B * py) + C) / D).astype(np.float32)         # float pl_distance = std::abs(...) / D(i)
        ok = box & ~(dist > np.float32(3))
        side1 = (lx1 - px) * (lx1 - px) + (ly1 - py) * (ly1 - py)
        side2 = (lx2 - px) * (lx2 - px) + (ly2 - py) * (ly2 - py)
        line_side = D * D
        ok &= (side1 <= 9) | (side2 <= 9) | ((side1 < line_side + side2) & (side2 < line_side + side1))
        rel.append({int(j): float(dist[j]) for j in np.nonzero(ok)[0]})
    return rel


def filter_stereo_matches(feat_l, feat_r, matches, min_x_diff, max_x_diff, max_y_diff):
    """Frame::AddRightFeatures, src/frame.cc:141-155: keep (q, t) with min < |xl - xr| < max and |yl - yr| <= max_y (float difference,
    compared as double).  matches: iterable of (queryIdx, trainIdx[, ...])."""
    out = []
    for m in matches:
        q, t = int(m[0]), int(m[1])
        dx = np.float64(np.abs(np.float32(feat_l[1, q]) - np.float32(feat_r[1, t])))
        dy = np.float64(np.abs(np.float32(feat_l[2, q]) - np.float32(feat_r[2, t])))
        if dx > min_x_diff and dx < max_x_diff and dy <= max_y_diff:
            out.append((q, t))
    return out


def match_lines(rel0, rel1, point_matches, n0, n1):
    """MatchLines, src/line_processor.cc:122-187.  rel0 / rel1 from assign_points_to_lines, point_matches list of (queryIdx, trainIdx).
    Returns line_matches [len(rel0)] (index into the second image's lines, -1 = none)."""
    L0, L1 = len(rel0), len(rel1)
    out = [-1] * L0
    if n0 == 0 or n1 == 0 or L0 == 0 or L1 == 0:
        return out
    a0 = [[] for _ in range(n0)]
    a1 = [[] for _ in range(n1)]
    for i, r in enumerate(rel0):
        for j in r:
            a0[j].append(i)
    for i, r in enumerate(rel1):
        for j in r:
            a1[j].append(i)
    mm = np.zeros((L0, L1), dtype=np.int32)
    for q, t in point_matches:
        for l0 in a0[q]:
            for l1 in a1[t]:
                mm[l0, l1] += 1
    row_loc = mm.argmax(axis=1)                   # Eigen maxCoeff(&index): first maximum
    for j in range(L1):
        loc = int(mm[:, j].argmax())
        v = int(mm[loc, j])
        if v < 2 or row_loc[loc] != j:
            continue
        score = np.float32(v * v) / np.float32(min(len(rel0[loc]), len(rel1[j])))
        if np.float64(score) < 0.8:                # float score compared with the double literal
            continue
        out[loc] = j
    return out
