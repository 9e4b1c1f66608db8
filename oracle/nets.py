"""Oracle restatement of the five model graphs on torch-CPU.  TEST INFRASTRUCTURE.

Each function restates, layer by layer, what the reference's TensorRT engines
compute for the corresponding shipped ONNX file (the graphs are the arithmetic
spec; the engines are built from them at src/super_point.cpp:18-85,
src/plnet.cpp:24-196, src/light_glue.cpp:24-118, src/super_glue.cpp:26-130 and
executed at src/super_point.cpp:133, src/plnet.cpp:233,510,
src/light_glue.cpp:159, src/super_glue.cpp:185).  Pinned against the reference's
ONNX files executed by OpenCV DNN (an external runtime; tests/golden/cv2dnn_*.npz,
tests/test_oracle_cv2dnn.py) and by the node-by-node interpreter
tools/onnx_interp.py (tests/golden/g*.npz, tests/test_oracle_golden.py; live in
the authoring container: tests/test_oracle_vs_onnx.py).

`emul=True` rounds both operands of every conv / matmul to fp16 and accumulates
in fp32: what the tcgen05 kernels compute up to summation order (SURVEY.md §8c
"emul" mode).  Weights in weights/*.afw are already fp16-representable.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F


def _r(x, emul):
    return x.to(torch.float16).to(torch.float32) if emul else x


def _attend(scores, v_of, emul):
    """softmax(scores) . V for one attention, `v_of(p)` = the P.V product for a probability operand p (already fp16-rounded when emul).
    emul == "fused" models WHERE the fused attention kernel (csrc/tc_attn.cuh) rounds: it hands fp16(exp(s - max)) to the tensor core and
    applies 1/sum to the fp32 product, whereas plain emul (and a framework executing the graph node by node) rounds the NORMALISED
    probabilities.  Both are fp16-operand realisations with the same relative rounding (2^-11), but the rounding errors are uncorrelated,
    and 18 stacked attentions amplify that into ~1e-2 on the match probability of ambiguous keypoints -- the same class of drift as
    fp32 vs emul (tools/experiments/r02_attention_rounding_drift.py has the numbers)."""
    if emul == "fused":
        e = torch.exp(scores - scores.amax(dim=-1, keepdim=True))
        return v_of(_r(e, True)) / e.sum(dim=-1, keepdim=True)
    return v_of(_r(torch.softmax(scores, dim=-1), emul))


def _t(a):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))


class W:
    """Weight view with a name prefix."""

    def __init__(self, w, prefix):
        self.w, self.p = w, prefix

    def __call__(self, name):
        return _t(self.w[self.p + name])

    def has(self, name):
        return (self.p + name) in self.w


def conv(x, w, name, emul, relu=True, pad=None):
    wt = w(name + ".weight")
    b = w(name + ".bias") if w.has(name + ".bias") else None
    if pad is None:
        pad = wt.shape[-1] // 2
    y = F.conv2d(_r(x, emul), _r(wt, emul), b, padding=pad)
    return F.relu(y) if relu else y


# ------------------------------------------------------------------------------------------------
# G1: SuperPoint  (output/superpoint_v1_sim_int32.onnx; also G2's /backbone/point_detector/*)
# ------------------------------------------------------------------------------------------------
def simple_nms(s):
    """In-graph NMS, radius 4 (G1 nodes 86..scores).  s: [B,H,W] float32."""
    def mp(x):
        return F.max_pool2d(x.unsqueeze(1), 9, 1, 4).squeeze(1)
    zeros = torch.zeros_like(s)
    m = s == mp(s)
    for _ in range(2):
        supp = mp(m.float()) > 0
        ss = torch.where(supp, zeros, s)
        m = m | ((ss == mp(ss)) & (~supp))
    return torch.where(m, s, zeros)


def superpoint_trunk(x, w, emul=False, keep=None):
    """VGG encoder; returns (relu_1 [64@512], relu_3 [64@256], relu_5 [128@128], relu_7 [128@64])."""
    a = conv(x, w, "conv1a", emul)
    r1 = conv(a, w, "conv1b", emul)
    a = F.max_pool2d(r1, 2, 2)
    a = conv(a, w, "conv2a", emul)
    r3 = conv(a, w, "conv2b", emul)
    a = F.max_pool2d(r3, 2, 2)
    a = conv(a, w, "conv3a", emul)
    r5 = conv(a, w, "conv3b", emul)
    a = F.max_pool2d(r5, 2, 2)
    a = conv(a, w, "conv4a", emul)
    r7 = conv(a, w, "conv4b", emul)
    if keep is not None:
        keep.update(relu_1=r1, relu_3=r3, relu_5=r5, relu_7=r7)
    return r1, r3, r5, r7


def superpoint_heads(r7, w, emul=False, keep=None):
    cpa = conv(r7, w, "convPa", emul)
    logits = conv(cpa, w, "convPb", emul, relu=False)            # [B,65,64,64]
    prob = torch.softmax(logits, dim=1)[:, :64]                 # drop dustbin
    b, _, hc, wc = prob.shape
    heat = prob.permute(0, 2, 3, 1).reshape(b, hc, wc, 8, 8).permute(0, 1, 3, 2, 4).reshape(b, hc * 8, wc * 8)
    scores = simple_nms(heat)
    cda = conv(r7, w, "convDa", emul)
    d = conv(cda, w, "convDb", emul, relu=False)                # [B,256,64,64]
    nrm = torch.sqrt((d * d).sum(dim=1, keepdim=True)).clamp_min(1e-12)
    desc = d / nrm
    if keep is not None:
        keep.update(logits=logits, heat=heat, desc_raw=d)
    return scores, desc


def superpoint_forward(x, weights, prefix="sp.", emul=False, keep=None):
    """x: [B,1,512,512] float32 in [0,1].  Returns scores [B,512,512] (post-NMS), descriptors [B,256,64,64]."""
    w = W(weights, prefix)
    with torch.no_grad():
        x = _t(x)
        _, _, _, r7 = superpoint_trunk(x, w, emul, keep)
        return superpoint_heads(r7, w, emul, keep)


# ------------------------------------------------------------------------------------------------
# G2: PLNet stage 0  (output/plnet_s0.onnx)
# ------------------------------------------------------------------------------------------------
def _up2(x):
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)   # Resize nearest/asymmetric/floor x2


def hourglass(x, w, s, emul):
    r = conv(x, w, s + "conv1a", emul)
    r1 = conv(r, w, s + "conv1b", emul)
    r = conv(F.max_pool2d(r1, 2, 2), w, s + "conv2a", emul)
    r3 = conv(r, w, s + "conv2b", emul)
    r = conv(F.max_pool2d(r3, 2, 2), w, s + "conv3a", emul)
    r5 = conv(r, w, s + "conv3b", emul)
    r = conv(F.max_pool2d(r5, 2, 2), w, s + "conv4a", emul)
    r7 = conv(r, w, s + "conv4b", emul)
    r = conv(F.max_pool2d(r7, 2, 2), w, s + "conv5a", emul)
    r9 = conv(r, w, s + "conv5b", emul)
    u = r9
    for lvl, skip in ((1, r7), (2, r5), (3, r3), (4, r1)):
        d = conv(_up2(u), w, s + "deconv%d" % lvl, emul)
        a = conv(skip, w, s + "conv%da_up" % (5 - lvl), emul)
        u = conv(torch.cat([d, a], dim=1), w, s + "conv%db_up" % (5 - lvl), emul)
    return u


def hafm_decode(heads9, joff_unused=None):
    """In-graph HAFM decode + junction top-K + association (G2 nodes 279-530).

    heads9: [1,9,128,128] raw head outputs (md 0-2, dis 3, res 4, jloc 5-6, joff 7-8).
    Returns dict with lines_pred [49152,4], juncs_pred [300,2], iskeep, idx_junc_to_end_min/max [49152] (float32).
    """
    assert heads9.shape[0] == 1
    md = torch.sigmoid(heads9[:, 0:3])
    dis = torch.sigmoid(heads9[:, 3:4])
    res = torch.sigmoid(heads9[:, 4:5])
    jloc = torch.softmax(heads9[:, 5:7], dim=1)[:, 1:]
    joff = torch.sigmoid(heads9[:, 7:9]) - 0.5
    h, wd = heads9.shape[2], heads9.shape[3]
    ys = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1).expand(1, 1, h, wd)
    xs = torch.arange(wd, dtype=torch.float32).view(1, 1, 1, wd).expand(1, 1, h, wd)
    sign = torch.tensor([-1.0, 0.0, 1.0]).view(1, 3, 1, 1)
    d = torch.clamp(dis + res * sign, 0.0, 1.0)                 # [1,3,H,W]
    pi = torch.tensor(3.1415927410125732)
    th = ((md[:, 0:1] - 0.5) * pi) * 2.0
    a = torch.tan((md[:, 1:2] * pi) / 2.0)
    b = torch.tan(((-md[:, 2:3]) * pi) / 2.0)
    c, s = torch.cos(th), torch.sin(th)
    x1 = torch.clamp(((c - s * a) * d) * 2.0 + xs, 0.0, float(wd - 1))
    y1 = torch.clamp(((s + c * a) * d) * 2.0 + ys, 0.0, float(h - 1))
    x2 = torch.clamp(((c - s * b) * d) * 2.0 + xs, 0.0, float(wd - 1))
    y2 = torch.clamp(((s + c * b) * d) * 2.0 + ys, 0.0, float(h - 1))
    lines = torch.stack([x1, y1, x2, y2], dim=-1).reshape(-1, 4)   # candidate-major, then y, then x
    return dict(lines_pred=lines, jloc=jloc, joff=joff, **junctions_and_association(lines, jloc, joff))


def junctions_and_association(lines, jloc, joff, topk=300):
    h, wd = jloc.shape[2], jloc.shape[3]
    keep = (jloc == F.max_pool2d(jloc, 3, 1, 1)).float()
    flat = (jloc * keep).reshape(-1)
    order = np.argsort(-flat.numpy(), kind="stable")[:topk]       # value desc, index asc (ONNX TopK)
    idx = torch.from_numpy(order.astype(np.int64))
    jo = joff.reshape(2, -1)
    jy = torch.div(idx, wd, rounding_mode="trunc").float() + jo[1][idx] + 0.5
    jx = torch.remainder(idx, wd).float() + jo[0][idx] + 0.5
    juncs = torch.stack([jx, jy], dim=1)                          # [300,2]
    d1 = ((lines[None, :, 0:2] - juncs[:, None, :]) ** 2).sum(-1)  # [300,L]
    d2 = ((lines[None, :, 2:4] - juncs[:, None, :]) ** 2).sum(-1)
    m1, m2 = d1.amin(0), d2.amin(0)
    i1 = torch.from_numpy(np.argmin(d1.numpy(), axis=0))          # first index on ties
    i2 = torch.from_numpy(np.argmin(d2.numpy(), axis=0))
    imin, imax = torch.minimum(i1, i2), torch.maximum(i1, i2)
    iskeep = (imin < imax) & (m1 < 10.0) & (m2 < 10.0)
    return dict(juncs_pred=juncs, iskeep=iskeep.float(), idx_junc_to_end_min=imin.float(),
                idx_junc_to_end_max=imax.float(), junc_topk_idx=idx)


def plnet_s0_forward(x, weights, emul=False, keep=None):
    """x: [1,1,512,512].  Returns the 10 outputs of plnet_s0.onnx (names as src/plnet.cpp:453-462)."""
    w = W(weights, "plnet.")
    wpd = W(weights, "plnet.pd.")
    with torch.no_grad():
        x = _t(x)
        r1, r3, r5, r7 = superpoint_trunk(x, wpd, emul, keep)
        scores, desc = superpoint_heads(r7, wpd, emul, keep)
        a = conv(r1, w, "conv1a", emul)
        a = conv(a, w, "conv1b", emul)
        a = torch.cat([F.max_pool2d(a, 2, 2), r3], dim=1)             # 32 | 64
        a = conv(a, w, "conv2a", emul)
        a = conv(a, w, "conv2b", emul)
        a = torch.cat([F.max_pool2d(a, 2, 2), r5], dim=1)             # 128 | 128
        a = hourglass(a, w, "stack1.", emul)
        a = hourglass(a, w, "stack2.", emul)
        f = conv(a, w, "fc2", emul, relu=False)                       # [1,256,128,128], no ReLU
        heads = []
        for i in range(5):
            hmid = conv(f, w, "score1.heads.%d.0" % i, emul)
            heads.append(conv(hmid, w, "score1.heads.%d.2" % i, emul, relu=False))
        heads9 = torch.cat(heads, dim=1)
        loi = conv(f, w, "fc1", emul, relu=False)
        thin = conv(f, w, "fc3", emul, relu=False)
        aux = conv(f, w, "fc4", emul, relu=False)
        dec = hafm_decode(heads9)
        if keep is not None:
            keep.update(heads9=heads9, fc2=f, stack_out=a)
        out = dict(iskeep=dec["iskeep"], idx_junc_to_end_min=dec["idx_junc_to_end_min"],
                   idx_junc_to_end_max=dec["idx_junc_to_end_max"], juncs_pred=dec["juncs_pred"],
                   lines_pred=dec["lines_pred"], loi_features=loi, loi_features_thin=thin,
                   loi_features_aux=aux, scores=scores, descriptors=desc)
        return out


# ------------------------------------------------------------------------------------------------
# G3: PLNet stage 1  (output/plnet_s1.onnx) -- HAWP-style line verification
# ------------------------------------------------------------------------------------------------
def bilinear_loi(feat, px, py):
    """feat [C,H,W]; px,py any shape.  Sampler of SURVEY.md App. B5 (weights from clamped integer coords)."""
    c, h, wd = feat.shape
    px = px - 0.5
    py = py - 0.5
    x0 = torch.floor(px).clamp(0, wd - 1)
    y0 = torch.floor(py).clamp(0, h - 1)
    x1 = (x0 + 1).clamp(0, wd - 1)
    y1 = (y0 + 1).clamp(0, h - 1)
    x0l, y0l, x1l, y1l = x0.long(), y0.long(), x1.long(), y1.long()
    f = feat
    return (f[:, y0l, x0l] * ((y1 - py) * (x1 - px)) + f[:, y1l, x0l] * ((py - y0) * (x1 - px))
            + f[:, y0l, x1l] * ((y1 - py) * (px - x0)) + f[:, y1l, x1l] * ((py - y0) * (px - x0)))


def plnet_s1_forward(juncs_pred, lines_pred, idx_lines_for_junctions, inverse, iskeep_index,
                     loi_features, loi_features_thin, loi_features_aux, weights, emul=False, keep=None):
    """All index inputs arrive as float32 arrays (src/plnet.cpp:494-507).  Returns lines_adjusted [U,4], scores_line [U]."""
    w = W(weights, "plnet.s1.")
    with torch.no_grad():
        juncs = _t(juncs_pred).float()
        lines = _t(lines_pred).float()
        pairs = _t(idx_lines_for_junctions).long().reshape(-1, 2)
        inv = _t(inverse).long().reshape(-1)
        kidx = _t(iskeep_index).long().reshape(-1)
        loi = _t(loi_features).float()[0]
        thin = _t(loi_features_thin).float()[0]
        aux = _t(loi_features_aux).float()[0]
        u = pairs.shape[0]
        adj = torch.cat([juncs[pairs[:, 0]], juncs[pairs[:, 1]]], dim=1)        # [U,4]
        # first kept proposal of each unique id (ScatterElements over reversed positions: last writer wins)
        first = np.full(u, -1, dtype=np.int64)
        invn = inv.numpy()
        pos = np.arange(len(invn), dtype=np.int64)
        first[invn[::-1]] = pos[::-1]
        orig = lines[kidx[torch.from_numpy(first)]]                             # [U,4]
        fe1 = bilinear_loi(loi, adj[:, 0], adj[:, 1]).t()                       # [U,128]
        fe2 = bilinear_loi(loi, adj[:, 2], adj[:, 3]).t()
        t = w("tspan").view(1, -1)
        tc = w("tspan_c").view(1, -1)

        def span(l4):
            sx = l4[:, 0:1] * t + l4[:, 2:3] * tc
            sy = l4[:, 1:2] * t + l4[:, 3:4] * tc
            return sx, sy
        sx, sy = span(adj)
        f_thin = bilinear_loi(thin, sx, sy).permute(1, 0, 2).reshape(u, -1)     # [U,4*30] channel-major
        sx, sy = span(orig)
        f_aux = bilinear_loi(aux, sx, sy).permute(1, 0, 2).reshape(u, -1)
        feat = torch.cat([fe1, fe2, f_thin, f_aux], dim=1)                      # [U,496]

        def lin(x, name, relu):
            y = F.linear(_r(x, emul), _r(w(name + ".weight"), emul), w(name + ".bias"))
            return F.relu(y) if relu else y
        h1 = lin(feat, "fc2.0", True)
        h1 = lin(h1, "fc2.2", True)
        h1 = lin(h1, "fc2.4", False)
        h2 = lin(torch.cat([f_thin, f_aux], dim=1), "fc2_res.0", True)
        logits = lin(h1 + h2, "fc2_head", False)
        score = torch.softmax(logits, dim=-1)[:, 1]
        if keep is not None:
            keep.update(feat=feat, orig=orig, logits=logits)
        return adj, score


# ------------------------------------------------------------------------------------------------
# G4: LightGlue  (output/superpoint_lightglue.onnx) -- 9 layers, no pruning / early exit
# ------------------------------------------------------------------------------------------------
def _rot_half(x):
    x = x.unflatten(-1, (-1, 2))
    x1, x2 = x.unbind(-1)
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def lightglue_forward(kpts0, kpts1, desc0, desc1, weights, emul=False, keep=None, n_layers=9):
    """kpts [N,2] (already normalised by PointMatcher::NormalizeKeypoints), desc [N,256] row-major.
    Returns log-assignment scores [N0,N1] (no dustbin)  (SURVEY.md App. B6)."""
    w = W(weights, "lg.")

    def lin(x, name):
        y = torch.matmul(_r(x, emul), _r(w(name + ".weight"), emul).t())
        return y + w(name + ".bias") if w.has(name + ".bias") else y

    def mm(a, b):
        return torch.matmul(_r(a, emul), _r(b, emul))

    with torch.no_grad():
        k0, k1 = _t(kpts0).float(), _t(kpts1).float()
        x0, x1 = _t(desc0).float(), _t(desc1).float()
        enc = []
        for k in (k0, k1):
            f = torch.matmul(_r(k, emul), _r(w("posenc.Wr.weight"), emul).t())   # [N,32]
            enc.append((torch.cos(f).repeat_interleave(2, dim=-1), torch.sin(f).repeat_interleave(2, dim=-1)))
        sc = 64 ** -0.25

        def ffn(x, msg, p):
            hcat = torch.cat([x, msg], dim=-1)
            hh = lin(hcat, p + "ffn.0")
            hh = F.layer_norm(hh, (512,), w(p + "ffn.1.weight"), w(p + "ffn.1.bias"), eps=1e-5)
            hh = F.gelu(hh)
            return x + lin(hh, p + "ffn.3")

        def heads(t):                     # [N,256] -> [4,N,64]
            return t.reshape(-1, 4, 64).permute(1, 0, 2)

        for l in range(n_layers):
            p = "transformers.%d.self_attn." % l
            new = []
            for x, (c, s) in ((x0, enc[0]), (x1, enc[1])):
                qkv = lin(x, p + "Wqkv").reshape(-1, 4, 64, 3)                   # feature = h*192 + d*3 + s
                q, k, v = qkv[..., 0].permute(1, 0, 2), qkv[..., 1].permute(1, 0, 2), qkv[..., 2].permute(1, 0, 2)
                q = q * c + _rot_half(q) * s
                k = k * c + _rot_half(k) * s
                ctx = _attend(mm(q * sc, (k * sc).transpose(-1, -2)), lambda pr, v=v: torch.matmul(pr, _r(v, emul)), emul).permute(1, 0, 2).reshape(-1, 256)
                msg = lin(ctx, p + "out_proj")
                new.append(ffn(x, msg, p))
            x0, x1 = new
            p = "transformers.%d.cross_attn." % l
            qk0, qk1 = heads(lin(x0, p + "to_qk")), heads(lin(x1, p + "to_qk"))
            v0, v1 = heads(lin(x0, p + "to_v")), heads(lin(x1, p + "to_v"))
            s01 = mm(qk0 * sc, (qk1 * sc).transpose(-1, -2))                    # [4,N0,N1]
            s10 = mm(qk1 * sc, (qk0 * sc).transpose(-1, -2))
            m0 = _attend(s01, lambda pr: torch.matmul(pr, _r(v1, emul)), emul).permute(1, 0, 2).reshape(-1, 256)
            m1 = _attend(s10, lambda pr: torch.matmul(pr, _r(v0, emul)), emul).permute(1, 0, 2).reshape(-1, 256)
            m0, m1 = lin(m0, p + "to_out"), lin(m1, p + "to_out")
            x0, x1 = ffn(x0, m0, p), ffn(x1, m1, p)
            if keep is not None:
                keep["x0_l%d" % l] = x0
                keep["x1_l%d" % l] = x1
        p = "log_assignment.8."
        md0 = lin(x0, p + "final_proj") / (256 ** 0.25)
        md1 = lin(x1, p + "final_proj") / (256 ** 0.25)
        sim = mm(md0, md1.t())
        z0 = lin(x0, p + "matchability")
        z1 = lin(x1, p + "matchability")
        scores = (torch.log_softmax(sim, dim=1) + torch.log_softmax(sim, dim=0)
                  + F.logsigmoid(z0) + F.logsigmoid(z1).t())
        if keep is not None:
            keep.update(sim=sim, z0=z0, z1=z1)
        return scores


# ------------------------------------------------------------------------------------------------
# G5: SuperGlue  (output/superglue_{indoor,outdoor}_sim_int32.onnx)
# ------------------------------------------------------------------------------------------------
def log_sinkhorn(z, log_mu, log_nu, iters):
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(z + v.unsqueeze(0), dim=1)
        v = log_nu - torch.logsumexp(z + u.unsqueeze(1), dim=0)
    return z + u.unsqueeze(1) + v.unsqueeze(0)


def superglue_forward(kpts0, scores0, desc0, kpts1, scores1, desc1, weights, emul=False, keep=None, iters=100):
    """kpts [N,2] normalised, scores [N], desc [256,N] channel-major (src/super_glue.cpp:199-246).
    Returns the [N0+1, N1+1] log-OT score matrix."""
    w = W(weights, "sg.")

    def c1(x, name, relu=False):          # Conv1d k=1 on [C,N]
        y = torch.matmul(_r(w(name + ".weight"), emul), _r(x, emul)) + w(name + ".bias").unsqueeze(1)
        return F.relu(y) if relu else y

    def mm(a, b):
        return torch.matmul(_r(a, emul), _r(b, emul))

    with torch.no_grad():
        d = []
        for k, s, de in ((kpts0, scores0, desc0), (kpts1, scores1, desc1)):
            k, s, de = _t(k).float(), _t(s).float(), _t(de).float()
            x = torch.cat([k.t(), s.view(1, -1)], dim=0)                       # [3,N]
            for i in (0, 3, 6, 9):
                x = c1(x, "kenc.encoder.%d" % i, relu=True)
            x = c1(x, "kenc.encoder.12")
            d.append(de + x)
        d0, d1 = d

        def attn(x, src, p):
            q, k, v = c1(x, p + "proj.0"), c1(src, p + "proj.1"), c1(src, p + "proj.2")
            n, m = q.shape[1], k.shape[1]
            q, k, v = q.view(64, 4, n), k.view(64, 4, m), v.view(64, 4, m)     # channel c -> (d=c//4, h=c%4)
            sc = torch.einsum("dhn,dhm->hnm", _r(q, emul), _r(k, emul)) / 8.0
            o = _attend(sc, lambda pr: torch.einsum("hnm,dhm->hnd", pr, _r(v, emul)), emul).permute(2, 0, 1).reshape(256, n)
            return c1(o, p + "merge")

        for l in range(18):
            p = "gnn.layers.%d." % l
            if l % 2 == 0:
                s0, s1 = d0, d1
            else:
                s0, s1 = d1, d0
            m0, m1 = attn(d0, s0, p + "attn."), attn(d1, s1, p + "attn.")

            def mlp(x, m):
                hh = c1(torch.cat([x, m], dim=0), p + "mlp.0", relu=True)
                return c1(hh, p + "mlp.3")
            d0, d1 = d0 + mlp(d0, m0), d1 + mlp(d1, m1)
        f0, f1 = c1(d0, "final_proj"), c1(d1, "final_proj")
        s = mm(f0.t(), f1) / 16.0
        m, n = s.shape
        alpha = w("bin_score").reshape(())
        z = torch.cat([torch.cat([s, alpha.expand(m, 1)], dim=1), alpha.expand(1, n + 1)], dim=0)
        norm = -math.log(m + n)
        norm_t = -torch.log(torch.tensor(float(m + n)))
        log_mu = torch.cat([norm_t.expand(m), (torch.log(torch.tensor(float(n))) + norm_t).view(1)])
        log_nu = torch.cat([norm_t.expand(n), (torch.log(torch.tensor(float(m))) + norm_t).view(1)])
        zz = log_sinkhorn(z, log_mu, log_nu, iters)
        if keep is not None:
            keep.update(couplings=z, sim=s)
        return zz - norm_t
