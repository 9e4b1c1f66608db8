"""ctypes binding of include/airfe_c.h.  Fails loudly if libairfe.so is missing: there is no fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libairfe.so")


class AirfeError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AirfeError("libairfe.so not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                             "there is no CPU fallback")
        from . import build as _build
        if not _build.is_current():
            # sources changed after the last build (or the stamp is missing): rebuild when a compiler is here, else refuse -- never run a
            # library that does not correspond to the sources next to it
            if os.path.exists(_build.NVCC):
                import fcntl
                with open(LIB_PATH + ".lock", "w") as lk:      # several ranks may get here at once: one builds, the others wait and re-check
                    fcntl.flock(lk, fcntl.LOCK_EX)
                    if not _build.is_current():
                        _build.build()
            else:
                raise AirfeError("libairfe.so is stale with respect to airslam_b200/csrc (no nvcc here to rebuild it)")
        _lib = C.CDLL(LIB_PATH)
        _lib.airfe_last_error.restype = C.c_char_p
        _lib.airfe_last_error.argtypes = [C.c_void_p]
        _declare(_lib)
    return _lib


def check(rc, ctx=None):
    if rc != 0:
        raise AirfeError("airfe error %d: %s" % (rc, lib().airfe_last_error(ctx).decode()))


vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_longlong, C.c_void_p


def _declare(L):
    L.airfe_op_tc_gemm.argtypes = [vp, i32, i32, i32, i32, i64, i64, i64,
                                   vp, i32, i32, i64, i64, i32, i32,
                                   i32, i32, i32, f32p, i32, i32,
                                   vp, i64, i64, i64, i32, i32, i32, i32, vp]
    L.airfe_op_tc_gemm.restype = i32
    L.airfe_op_conv3x3.argtypes = [vp, i32, i32, i32, i32, i64, vp, f32p, i32, i32, i32, vp, i64, vp, i64, vp]
    L.airfe_op_conv3x3.restype = i32


class Config(C.Structure):
    _fields_ = [("weights_dir", C.c_char_p), ("max_batch", i32), ("max_keypoints", i32), ("keypoint_threshold", C.c_float),
                ("remove_borders", i32), ("line_threshold", C.c_float), ("line_length_threshold", C.c_float),
                ("image_width", i32), ("image_height", i32), ("enable_superpoint", i32), ("enable_plnet", i32),
                ("enable_lightglue", i32), ("enable_superglue", i32)]


NET_SUPERPOINT, NET_PLNET = 0, 1
MATCHER_LIGHTGLUE, MATCHER_SUPERGLUE = 0, 1
WEIGHTS_DIR = os.path.join(os.path.dirname(HERE), "weights")


def _declare_frame(L):
    L.airfe_alloc_pinned.argtypes = [i64]
    L.airfe_alloc_pinned.restype = vp
    L.airfe_free_pinned.argtypes = [vp]
    L.airfe_default_config.argtypes = [C.POINTER(Config)]
    L.airfe_create.argtypes = [C.POINTER(Config), i32, C.POINTER(vp)]
    L.airfe_create.restype = i32
    L.airfe_destroy.argtypes = [vp]
    L.airfe_debug_conv_trace.argtypes = [vp]
    L.airfe_debug_conv_trace.restype = None
    L.airfe_debug_match_trace.argtypes = [vp]
    L.airfe_debug_match_trace.restype = None
    L.airfe_stream.argtypes = [vp]
    L.airfe_stream.restype = vp
    L.airfe_detect_batch.argtypes = [vp, i32, i32, vp, i32, i32, i32, i64, vp, i32, vp, vp, i32, vp, vp, i32, vp]
    L.airfe_detect_batch.restype = i32
    L.airfe_debug_read.argtypes = [vp, i32, C.c_char_p, i32, vp, i64]
    L.airfe_debug_read.restype = i64


_declare_base = _declare


def _declare(L):  # noqa: F811
    _declare_base(L)
    _declare_frame(L)
    if hasattr(L, "airfe_match_batch"):
        _declare_match(L)


def _declare_match(L):
    L.airfe_match_batch.argtypes = [vp, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp]
    L.airfe_match_batch.restype = i32
    L.airfe_detect_match_stereo_batch.argtypes = [vp, i32, i32, i32, vp, vp, i32, i32, i32, i64, vp, i32, vp, vp, i32, vp, vp, i32, vp,
                                                  vp, vp, vp, i32, vp]
    L.airfe_detect_match_stereo_batch.restype = i32
    L.airfe_superglue_batch.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, i32]
    L.airfe_superglue_batch.restype = i32
    L.airfe_stereo_device.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, i64, i32, i32]
    L.airfe_stereo_device.restype = i32
    L.airfe_profile_stereo.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, i64, i32, i32, vp, i64]
    L.airfe_profile_stereo.restype = i64
    L.airfe_stereo_cost.argtypes = [vp, i32, i32, i32, i32, C.POINTER(C.c_double), C.POINTER(i32)]
    L.airfe_stereo_cost.restype = i32
    L.airfe_set_rectify_maps.argtypes = [vp, i32, vp, vp, i32, i32]
    L.airfe_set_rectify_maps.restype = i32
    L.airfe_set_rectify.argtypes = [vp, i32]
    L.airfe_set_rectify.restype = i32
    L.airfe_undistort.argtypes = [vp, i32, vp, i32, i32, i32, vp, i32]
    L.airfe_undistort.restype = i32
    L.airfe_stereo_line_assoc.argtypes = [vp, i32, C.c_double, C.c_double, C.c_double, i32, i32, vp, vp, vp, vp]
    L.airfe_stereo_line_assoc.restype = i32
    L.airfe_bow_load.argtypes = [vp, C.c_char_p]
    L.airfe_bow_load.restype = i32
    L.airfe_bow_transform.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    L.airfe_bow_transform.restype = i32
    L.airfe_kf_reserve.argtypes = [vp, i32, i32]
    L.airfe_kf_reserve.restype = i32
    L.airfe_kf_put.argtypes = [vp, i32, vp, i32]
    L.airfe_kf_put.restype = i32
    L.airfe_kf_size.argtypes = [vp]
    L.airfe_kf_size.restype = i32
    L.airfe_reloc_match.argtypes = [vp, i32, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32]
    L.airfe_reloc_match.restype = i32
    L.airfe_reloc_pick.argtypes = [i32, i32, vp, vp, vp]
    L.airfe_reloc_pick.restype = None


def pinned_array(shape, dtype):
    """numpy array backed by page-locked host memory (airfe_alloc_pinned); lives until the process exits."""
    import numpy as np
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib().airfe_alloc_pinned(max(n, 16))
    if not ptr:
        raise AirfeError("pinned allocation failed")
    buf = (C.c_char * max(n, 16)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class Context:
    """Owns one airfe_ctx.  numpy in / numpy out; every call goes through the C ABI with host buffers."""

    def __init__(self, device=0, **kw):
        import numpy as np  # noqa: F401
        L = lib()
        self.cfg = Config()
        L.airfe_default_config(C.byref(self.cfg))
        self._wd = WEIGHTS_DIR.encode()
        self.cfg.weights_dir = self._wd
        for k, v in kw.items():
            if not hasattr(self.cfg, k):
                raise AirfeError("unknown config field " + k)
            setattr(self.cfg, k, v)
        self.h = vp()
        check(L.airfe_create(C.byref(self.cfg), device, C.byref(self.h)))

    def _check(self, rc):
        check(rc, self.h)

    def close(self):
        if self.h:
            lib().airfe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return lib().airfe_stream(self.h)

    def detect_batch(self, net, images, lines=False, junctions=False, feat_cap=None, line_cap=4096, junc_cap=1024):
        """images: uint8 [B,H,W] contiguous.  Returns list of (feat [259,N], lines [L,4] float64 | None, junc [259,J] | None)."""
        import numpy as np
        images = np.ascontiguousarray(images, dtype=np.uint8)
        b, h, w = images.shape
        feat_cap = feat_cap or self.cfg.max_keypoints
        feat = np.zeros((b, feat_cap, 259), dtype=np.float32)
        n_feat = np.zeros(b, dtype=np.int32)
        ln = np.zeros((b, line_cap, 4), dtype=np.float64) if lines else None
        n_ln = np.zeros(b, dtype=np.int32)
        jn = np.zeros((b, junc_cap, 259), dtype=np.float32) if junctions else None
        n_jn = np.zeros(b, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(vp) if a is not None else None
        self._check(lib().airfe_detect_batch(self.h, net, b, p(images), w, h, w, h * w, p(feat), feat_cap, p(n_feat), p(ln), line_cap,
                                       p(n_ln), p(jn), junc_cap, p(n_jn)))
        out = []
        for i in range(b):
            out.append((feat[i, :n_feat[i]].T.copy(), ln[i, :n_ln[i]].copy() if lines else None,
                        jn[i, :n_jn[i]].T.copy() if junctions else None))
        return out

    def debug_read(self, net, name, index, dtype, shape):
        import numpy as np
        out = np.zeros(shape, dtype=dtype)
        n = lib().airfe_debug_read(self.h, net, name.encode(), index, out.ctypes.data_as(vp), out.nbytes)
        if n < 0:
            self._check(int(n))
        return out

    def match_batch(self, matcher, feats0, feats1, match_cap=1024):
        """feats0/feats1: lists of [259,N] float32 arrays (image-pixel keypoints).  Returns list of (idx [K,2] int32, score [K])."""
        import numpy as np
        p = len(feats0)
        cap = max(max(f.shape[1] for f in feats0), max(f.shape[1] for f in feats1), 1)
        f0 = np.zeros((p, cap, 259), dtype=np.float32)
        f1 = np.zeros((p, cap, 259), dtype=np.float32)
        n0 = np.array([f.shape[1] for f in feats0], dtype=np.int32)
        n1 = np.array([f.shape[1] for f in feats1], dtype=np.int32)
        for i in range(p):
            f0[i, :n0[i]] = feats0[i].T
            f1[i, :n1[i]] = feats1[i].T
        i0 = np.zeros((p, match_cap), dtype=np.int32)
        i1 = np.zeros((p, match_cap), dtype=np.int32)
        sc = np.zeros((p, match_cap), dtype=np.float32)
        nm = np.zeros(p, dtype=np.int32)
        q = lambda a: a.ctypes.data_as(vp)
        self._check(lib().airfe_match_batch(self.h, matcher, p, q(f0), q(n0), q(f1), q(n1), cap, q(i0), q(i1), q(sc), match_cap, q(nm)))
        return [(np.stack([i0[i, :nm[i]], i1[i, :nm[i]]], axis=1), sc[i, :nm[i]].copy()) for i in range(p)]

    def stereo_batch(self, net, matcher, left, right, lines=False, junctions=False, line_cap=2048, junc_cap=512, match_cap=1024, raw=False):
        """left/right: uint8 [P,H,W].  Returns per pair dict(feat_l, feat_r, lines_l, lines_r, junc, matches (idx, score)).
        Output buffers are allocated once per shape and reused (raw=True returns them unsliced: what a C caller sees)."""
        import numpy as np
        if not (left.dtype == np.uint8 and left.flags.c_contiguous):
            left = np.ascontiguousarray(left, dtype=np.uint8)
        if not (right.dtype == np.uint8 and right.flags.c_contiguous):
            right = np.ascontiguousarray(right, dtype=np.uint8)
        p, h, w = left.shape
        fc = self.cfg.max_keypoints
        key = (p, fc, bool(lines), bool(junctions), line_cap, junc_cap, match_cap)
        if not hasattr(self, "_sb_cache"):
            self._sb_cache = {}
        if key not in self._sb_cache:
            self._sb_cache[key] = dict(feat=pinned_array((2 * p, fc, 259), np.float32), nf=np.zeros(2 * p, dtype=np.int32),
                            ln=np.empty((2 * p, line_cap, 4), dtype=np.float64) if lines else None, nl=np.zeros(2 * p, dtype=np.int32),
                            jn=pinned_array((p, junc_cap, 259), np.float32) if junctions else None, nj=np.zeros(p, dtype=np.int32),
                            i0=np.empty((p, match_cap), dtype=np.int32), i1=np.empty((p, match_cap), dtype=np.int32),
                            sc=np.empty((p, match_cap), dtype=np.float32), nm=np.zeros(p, dtype=np.int32))
        b = self._sb_cache[key]
        feat, nf, ln, nl, jn, nj, i0, i1, sc, nm = (b[k] for k in ("feat", "nf", "ln", "nl", "jn", "nj", "i0", "i1", "sc", "nm"))
        q = lambda a: a.ctypes.data_as(vp) if a is not None else None
        self._check(lib().airfe_detect_match_stereo_batch(self.h, net, matcher, p, q(left), q(right), w, h, w, h * w, q(feat), fc, q(nf), q(ln), line_cap,
                                                    q(nl), q(jn), junc_cap, q(nj), q(i0), q(i1), q(sc), match_cap, q(nm)))
        if raw:
            return b
        out = []
        for i in range(p):
            out.append(dict(feat_l=feat[2 * i, :nf[2 * i]].T.copy(), feat_r=feat[2 * i + 1, :nf[2 * i + 1]].T.copy(),
                            lines_l=ln[2 * i, :nl[2 * i]].copy() if lines else None, lines_r=ln[2 * i + 1, :nl[2 * i + 1]].copy() if lines else None,
                            junc=jn[i, :nj[i]].T.copy() if junctions else None,
                            matches=(np.stack([i0[i, :nm[i]], i1[i, :nm[i]]], axis=1), sc[i, :nm[i]].copy())))
        return out

    def stereo_device(self, net, matcher, pairs, d_images_ptr, w, h, stride, img_stride, lines=True, junctions=True):
        self._check(lib().airfe_stereo_device(self.h, net, matcher, pairs, d_images_ptr, w, h, stride, img_stride, int(lines), int(junctions)))

    def stereo_cost(self, net, matcher, pairs, lines=True):
        fl, ln = C.c_double(0), i32(0)
        self._check(lib().airfe_stereo_cost(self.h, net, matcher, pairs, int(lines), C.byref(fl), C.byref(ln)))
        return fl.value, ln.value

    def profile_stereo(self, net, matcher, pairs, d_images_ptr, w, h, stride, img_stride, lines=True, junctions=True):
        """One profiled step; returns list of (name, flops, ms)."""
        buf = C.create_string_buffer(1 << 18)
        n = lib().airfe_profile_stereo(self.h, net, matcher, pairs, d_images_ptr, w, h, stride, img_stride, int(lines), int(junctions), buf, len(buf))
        if n < 0:
            self._check(int(n))
        out = []
        for line in buf.value.decode().splitlines():
            nm, fl, ms = line.split("\t")
            out.append((nm, float(fl), float(ms)))
        return out

    # ---- BoW quantisation (Database::FrameToBow) ----
    def bow_load(self, path=None):
        self._check(lib().airfe_bow_load(self.h, (path or os.path.join(WEIGHTS_DIR, "point_voc_L4.afw")).encode()))

    def bow_transform(self, feat):
        """feat [259, N] float32 -> (word_of_feature uint32 [N], [(word id, normalised weight), ...])."""
        import numpy as np
        a = np.ascontiguousarray(feat.T, dtype=np.float32)
        n = a.shape[0]
        words = np.zeros(max(n, 1), np.uint32); ids = np.zeros(max(n, 1), np.uint32); vals = np.zeros(max(n, 1), np.float64)
        nb = i32(0)
        self._check(lib().airfe_bow_transform(self.h, a.ctypes.data_as(vp), n, words.ctypes.data_as(vp), ids.ctypes.data_as(vp), vals.ctypes.data_as(vp), C.byref(nb)))
        return words[:n], list(zip(ids[:nb.value].tolist(), vals[:nb.value].tolist()))

    # ---- point <-> line association + stereo line matching on the last stereo call's device-resident results ----
    def stereo_line_assoc(self, pairs, min_x_diff, max_x_diff, max_y_diff, line_cap=256, rel_cap=32):
        """Returns per pair (rel_left, rel_right, line_matches): rel_* = list (one per line slot) of {point index: distance}."""
        import numpy as np
        rn = np.zeros((2 * pairs, line_cap), np.int32)
        ri = np.zeros((2 * pairs, line_cap, rel_cap), np.int32)
        rd = np.zeros((2 * pairs, line_cap, rel_cap), np.float32)
        lm = np.zeros((pairs, line_cap), np.int32)
        q = lambda a: a.ctypes.data_as(vp)
        self._check(lib().airfe_stereo_line_assoc(self.h, pairs, min_x_diff, max_x_diff, max_y_diff, line_cap, rel_cap, q(rn), q(ri), q(rd), q(lm)))
        return rn, ri, rd, lm

    # ---- rectification (Camera::UndistortImage) fused into the first kernel ----
    def set_rectify_maps(self, side, map_x, map_y):
        import numpy as np
        mx = np.ascontiguousarray(map_x, dtype=np.float32)
        my = np.ascontiguousarray(map_y, dtype=np.float32)
        self._check(lib().airfe_set_rectify_maps(self.h, side, mx.ctypes.data_as(vp), my.ctypes.data_as(vp), mx.shape[1], mx.shape[0]))

    def set_rectify(self, mode):
        self._check(lib().airfe_set_rectify(self.h, mode))

    def undistort(self, side, raw):
        import numpy as np
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        out = np.empty_like(raw)
        self._check(lib().airfe_undistort(self.h, side, raw.ctypes.data_as(vp), raw.shape[1], raw.shape[0], raw.shape[1], out.ctypes.data_as(vp), raw.shape[1]))
        return out

    # ---- device-resident keyframe features + batched candidate matching (config 5) ----
    def kf_reserve(self, n_keyframes, feat_cap):
        self._check(lib().airfe_kf_reserve(self.h, n_keyframes, feat_cap))

    def kf_put(self, slot, feat):
        """feat: [259, N] float32 (the class-surface layout) on the host."""
        import numpy as np
        a = np.ascontiguousarray(feat.T, dtype=np.float32)        # column-major 259 x N == row-major [N][259]
        self._check(lib().airfe_kf_put(self.h, slot, a.ctypes.data_as(vp), a.shape[0]))

    def kf_put_ptr(self, slot, ptr, n):
        """Raw pointer (host or device) to [n][259] float32."""
        self._check(lib().airfe_kf_put(self.h, slot, ptr, n))

    def reloc_match(self, matcher, query_feat_ptr, query_n, feat_cap, job_query, job_kf, want_matches=False, match_cap=1024):
        """query_feat_ptr: pointer (host or device) to [Q][feat_cap][259] float32; query_n: int32 [Q] (host).  Returns counts [J]
        (and per-job (idx [K,2], score [K]) when want_matches)."""
        import numpy as np
        qn = np.ascontiguousarray(query_n, dtype=np.int32)
        jq = np.ascontiguousarray(job_query, dtype=np.int32)
        jk = np.ascontiguousarray(job_kf, dtype=np.int32)
        J = len(jq)
        nm = np.zeros(max(J, 1), dtype=np.int32)
        q = lambda a: a.ctypes.data_as(vp)
        if want_matches:
            i0 = np.zeros((max(J, 1), match_cap), np.int32); i1 = np.zeros((max(J, 1), match_cap), np.int32); sc = np.zeros((max(J, 1), match_cap), np.float32)
            self._check(lib().airfe_reloc_match(self.h, matcher, query_feat_ptr, q(qn), len(qn), feat_cap, J, q(jq), q(jk), q(nm), q(i0), q(i1), q(sc), match_cap))
            return nm[:J], [(np.stack([i0[j, :nm[j]], i1[j, :nm[j]]], 1), sc[j, :nm[j]].copy()) for j in range(J)]
        self._check(lib().airfe_reloc_match(self.h, matcher, query_feat_ptr, q(qn), len(qn), feat_cap, J, q(jq), q(jk), q(nm), None, None, None, 0))
        return nm[:J]

    def superglue_batch(self, feats0, feats1):
        """Raw SuperGlue::infer outputs per pair: (indices0, indices1, mscores0, mscores1)."""
        import numpy as np
        p = len(feats0)
        cap = max(max(f.shape[1] for f in feats0), max(f.shape[1] for f in feats1), 1)
        f0 = np.zeros((p, cap, 259), dtype=np.float32)
        f1 = np.zeros((p, cap, 259), dtype=np.float32)
        n0 = np.array([f.shape[1] for f in feats0], dtype=np.int32)
        n1 = np.array([f.shape[1] for f in feats1], dtype=np.int32)
        for i in range(p):
            f0[i, :n0[i]] = feats0[i].T
            f1[i, :n1[i]] = feats1[i].T
        i0 = np.zeros((p, cap), dtype=np.int32); i1 = np.zeros((p, cap), dtype=np.int32)
        m0 = np.zeros((p, cap), dtype=np.float32); m1 = np.zeros((p, cap), dtype=np.float32)
        q = lambda a: a.ctypes.data_as(vp)
        self._check(lib().airfe_superglue_batch(self.h, p, q(f0), q(n0), q(f1), q(n1), cap, 0, q(i0), q(i1), q(m0), q(m1), cap))
        return [(i0[i, :n0[i]].copy(), i1[i, :n1[i]].copy(), m0[i, :n0[i]].copy(), m1[i, :n1[i]].copy()) for i in range(p)]


def reloc_pick(counts):
    """airfe_reloc_pick: winner per query under the rule of src/map_user.cc:370-373.  counts int32 [Q, C] (negative = absent)."""
    import numpy as np
    c = np.ascontiguousarray(counts, dtype=np.int32)
    best = np.zeros(c.shape[0], np.int32)
    cnt = np.zeros(c.shape[0], np.int32)
    lib().airfe_reloc_pick(c.shape[0], c.shape[1], c.ctypes.data_as(vp), best.ctypes.data_as(vp), cnt.ctypes.data_as(vp))
    return best, cnt
