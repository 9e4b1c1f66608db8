"""Seeded synthetic inputs for tests and bench.  Pure numpy, deterministic across machines.
TEST / BENCH INFRASTRUCTURE (no reference counterpart: the reference ships no data)."""
import numpy as np


def _blur(a, k):
    """Separable box blur applied twice (~triangular), reflect padding; integer-free float64 math."""
    if k <= 1:
        return a
    ker = np.ones(k) / k
    for _ in range(2):
        a = np.apply_along_axis(lambda r: np.convolve(np.pad(r, (k // 2, k - 1 - k // 2), mode="reflect"), ker, mode="valid"), 1, a)
        a = np.apply_along_axis(lambda r: np.convolve(np.pad(r, (k // 2, k - 1 - k // 2), mode="reflect"), ker, mode="valid"), 0, a)
    return a


def _draw_line(img, x0, y0, x1, y1, val, thick):
    n = int(max(abs(x1 - x0), abs(y1 - y0))) * 2 + 2
    t = np.linspace(0.0, 1.0, n)
    xs = np.rint(x0 + (x1 - x0) * t).astype(int)
    ys = np.rint(y0 + (y1 - y0) * t).astype(int)
    h, w = img.shape
    for dy in range(-(thick // 2), thick - thick // 2):
        for dx in range(-(thick // 2), thick - thick // 2):
            xx, yy = xs + dx, ys + dy
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            img[yy[ok], xx[ok]] = val


def scene(w, h, seed, margin=96):
    """Float64 image [h, w+margin] in 0..255 with texture, rectangles and line segments."""
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    ww = w + margin
    img = np.zeros((h, ww))
    for octave, k in enumerate((33, 17, 9, 5)):
        img += _blur(rs.uniform(0, 1, (h, ww)), k) * (0.5 ** octave) * 4.0
    img = (img - img.min()) / (img.max() - img.min()) * 140.0 + 50.0
    for _ in range(30):
        x0, y0 = rs.randint(0, ww - 20), rs.randint(0, h - 20)
        rw, rh = rs.randint(16, 120), rs.randint(16, 100)
        img[y0:y0 + rh, x0:x0 + rw] = rs.uniform(10, 245)
    for _ in range(40):
        x0, y0, x1, y1 = rs.uniform(0, ww), rs.uniform(0, h), rs.uniform(0, ww), rs.uniform(0, h)
        _draw_line(img, x0, y0, x1, y1, rs.choice([15.0, 240.0]), int(rs.randint(1, 4)))
    return img


def stereo_pair(w, h, seed, low_light=False):
    """Returns (left, right) uint8 [h,w]; right = left shifted by a per-pair disparity in [4,40] px + noise sigma 2."""
    rs = np.random.RandomState((seed * 2654435761 + 12345) & 0x7FFFFFFF)
    sc = scene(w, h, seed)
    disp = int(rs.randint(4, 41))
    left = sc[:, 8:8 + w].copy()
    right = sc[:, 8 + disp:8 + disp + w].copy()      # x_right = x_left - disp (rectified stereo geometry)
    if low_light:
        left, right = left * 0.25, right * 0.25
        left += rs.normal(0, 6.0, left.shape)
        right += rs.normal(0, 6.0, right.shape)
    else:
        left += rs.normal(0, 1.0, left.shape)
        right += rs.normal(0, 2.0, right.shape)
    to8 = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)
    return to8(left), to8(right), disp


def keypoint_set(n, w, h, seed, perturb_of=None):
    """Synthetic 259xN feature matrix: uniform keypoints, unit-norm Gaussian descriptors.  If perturb_of is a
    259xN matrix, returns a permuted, noised copy (planted correspondences) plus the permutation."""
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    if perturb_of is None:
        f = np.zeros((259, n), dtype=np.float32)
        f[0] = rs.uniform(0.01, 0.3, n)
        f[1] = rs.uniform(8, w - 8, n)
        f[2] = rs.uniform(8, h - 8, n)
        d = rs.normal(0, 1, (256, n))
        f[3:] = d / np.linalg.norm(d, axis=0, keepdims=True)
        return f
    assert n <= perturb_of.shape[1]
    perm = rs.permutation(perturb_of.shape[1])[:n]
    f = perturb_of[:, perm].copy()
    f[1] += rs.normal(0, 1.0, n) - 12.0
    f[2] += rs.normal(0, 1.0, n)
    d = f[3:] + rs.normal(0, 0.02, (256, n))
    f[3:] = d / np.linalg.norm(d, axis=0, keepdims=True)
    return f.astype(np.float32), perm
