"""GPU parity of the tcgen05 implicit-GEMM kernel (csrc/tc_gemm.cuh) against a plain fp32 torch reference with the
same fp16-rounded operands.  Tolerance: fp32 accumulation order only -> 1e-3 relative to the output scale."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from airslam_b200 import capi
    return capi.lib(), capi.check


def pack_conv_weight(w_oihw, c_pad):
    """OIHW fp32 -> [O][taps][c_pad] fp16 (K-major, tap-major)."""
    o, i, kh, kw = w_oihw.shape
    out = torch.zeros(o, kh * kw, c_pad, dtype=torch.float16, device=w_oihw.device)
    out[:, :, :i] = w_oihw.permute(0, 2, 3, 1).reshape(o, kh * kw, i).half()
    return out.reshape(o, kh * kw * c_pad).contiguous()


def run_conv(x_nhwc, w_oihw, bias, relu, block_n, tile, out_f32=False, out=None, out_ch_off=0, n_pad_rows=None):
    lib, check = _lib()
    b, h, w, c = x_nhwc.shape
    o, i, kh, kw = w_oihw.shape
    taps = kh * kw
    c_pad = (i + 63) // 64 * 64
    wp = pack_conv_weight(w_oihw, c_pad)
    if out is None:
        out = torch.zeros(b, h, w, o, dtype=torch.float32 if out_f32 else torch.float16, device=x_nhwc.device)
    c_tot = out.shape[-1]
    esz = out.element_size()
    optr = out.data_ptr() + out_ch_off * esz
    bptr = bias.data_ptr() if bias is not None else None
    tw, th, tb = tile
    check(lib.airfe_op_tc_gemm(x_nhwc.data_ptr(), i, w, h, b, x_nhwc.stride(2), x_nhwc.stride(1), x_nhwc.stride(0),
                               wp.data_ptr(), taps * c_pad, o, taps * c_pad, 0, 0, 0,
                               taps, c_pad, block_n, bptr, int(relu), int(out_f32),
                               optr, h * w * c_tot, w * c_tot, c_tot, o, tw, th, tb, None))
    torch.cuda.synchronize()
    return out


def ref_conv(x_nhwc, w_oihw, bias, relu):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w_oihw.half().float(), bias, padding=w_oihw.shape[-1] // 2)
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1)


def _mk(b, h, w, cin, cout, k, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(b, h, w, cin, device="cuda", generator=g)).half()
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) * (1.0 / (cin * k * k) ** 0.5)
    bias = torch.randn(cout, device="cuda", generator=g)
    return x, wt, bias


def _close(a, b, tol=2e-3):
    a, b = a.float(), b.float()
    scale = b.abs().max().item() + 1e-6
    err = (a - b).abs().max().item()
    assert err <= tol * scale, "max err %g (scale %g)" % (err, scale)


@pytest.mark.parametrize("cfg", [
    # b, h, w, cin, cout, k, block_n, tile
    (1, 1, 300, 128, 64, 1, 64, (128, 1, 1)),      # plain GEMM, ragged M
    (2, 32, 48, 64, 64, 3, 64, (16, 8, 1)),        # 3x3 conv, padding via TMA OOB
    (1, 16, 16, 128, 128, 3, 128, (16, 8, 1)),     # two K blocks per tap
    (1, 24, 40, 96, 32, 3, 32, (16, 8, 1)),        # C_in not a multiple of 64 (channel OOB fill)
    (3, 8, 8, 128, 128, 3, 128, (8, 8, 2)),        # 8x8 maps: two images per tile, odd batch
    (1, 64, 64, 128, 256, 3, 256, (16, 8, 1)),     # N = 256
    (2, 40, 56, 32, 32, 3, 32, (16, 8, 1)),        # C_in = 32
    (1, 128, 128, 256, 128, 1, 128, (16, 8, 1)),   # 1x1 conv, K = 256
])
def test_conv_matches_torch(cfg):
    b, h, w, cin, cout, k, bn, tile = cfg
    x, wt, bias = _mk(b, h, w, cin, cout, k, seed=hash(cfg) & 0xFFFF)
    y = run_conv(x, wt, bias, True, bn, tile)
    _close(y, ref_conv(x, wt, bias, True).half())


def test_fp32_out_ragged_n():
    x, wt, bias = _mk(1, 64, 64, 256, 65, 1, seed=5)
    out = torch.zeros(1, 64, 64, 80, dtype=torch.float32, device="cuda")   # pixel stride padded to the N tile
    y = run_conv(x, wt, bias, False, 80, (16, 8, 1), out_f32=True, out=out)
    _close(y[..., :65], ref_conv(x, wt, bias, False), tol=1e-3)
    assert float(y[..., 65:].abs().max()) == 0.0


def test_channel_offset_store_and_strided_input():
    """Concat by construction: two convs write disjoint channel ranges of one buffer; a third reads a channel slice."""
    x, wt, bias = _mk(1, 32, 32, 64, 64, 3, seed=9)
    buf = torch.zeros(1, 32, 32, 96, dtype=torch.float16, device="cuda")
    run_conv(x, wt[:32], bias[:32], True, 32, (16, 8, 1), out=buf, out_ch_off=0)
    run_conv(x, wt, bias, True, 64, (16, 8, 1), out=buf, out_ch_off=32)
    ref = ref_conv(x, wt, bias, True).half()
    _close(buf[..., :32], ref[..., :32])
    _close(buf[..., 32:], ref)
    # read the 64-channel slice [32:96) of the 96-channel buffer as a conv input
    x2 = buf[..., 32:]
    _, wt2, b2 = _mk(1, 32, 32, 64, 64, 3, seed=10)
    y2 = run_conv(x2, wt2, b2, False, 64, (16, 8, 1))
    _close(y2, ref_conv(x2.contiguous(), wt2, b2, False).half())


def test_batched_b_heads_qk():
    """Attention scores: per head h, S_h = Q_h K_h^T; heads are the batch dim of both operands."""
    lib, check = _lib()
    n0, n1 = 400, 380
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(n0, 256, device="cuda", generator=g).half()
    k = torch.randn(n1, 256, device="cuda", generator=g).half()
    n1p = (n1 + 63) // 64 * 64
    s = torch.zeros(4, n0, n1p, dtype=torch.float32, device="cuda")
    check(lib.airfe_op_tc_gemm(q.data_ptr(), 64, n0, 1, 4, 256, 256 * n0, 64,
                               k.data_ptr(), 64, n1, 256, 64, 4, 0,
                               1, 64, 128, None, 0, 1,
                               s.data_ptr(), n0 * n1p, 0, n1p, n1, 128, 1, 1, None))
    torch.cuda.synchronize()
    ref = torch.einsum("nhd,mhd->hnm", q.float().view(n0, 4, 64), k.float().view(n1, 4, 64))
    _close(s[:, :, :n1], ref, tol=1e-3)


def test_mn_major_b_pv():
    """O_h = P_h V_h with V stored [n1, 4*64] row-major (MN-major B operand, no transpose anywhere)."""
    lib, check = _lib()
    n0, n1 = 400, 380
    n1p = (n1 + 63) // 64 * 64
    g = torch.Generator(device="cuda").manual_seed(4)
    pm = torch.zeros(4, n0, n1p, device="cuda", dtype=torch.float16)
    pm[:, :, :n1] = torch.softmax(torch.randn(4, n0, n1, device="cuda", generator=g), dim=-1).half()
    v = torch.randn(n1, 256, device="cuda", generator=g).half()
    o = torch.zeros(n0, 256, dtype=torch.float16, device="cuda")
    check(lib.airfe_op_tc_gemm(pm.data_ptr(), n1p, n0, 1, 4, n1p, n0 * n1p, n0 * n1p,
                               v.data_ptr(), n1, 64, 256, 64, 4, 1,
                               1, n1p, 64, None, 0, 0,
                               o.data_ptr(), 64, 0, 256, 64, 128, 1, 1, None))
    torch.cuda.synchronize()
    ref = torch.einsum("hnm,mhd->nhd", pm[:, :, :n1].float(), v.float().view(n1, 4, 64)).reshape(n0, 256)
    _close(o, ref.half())


# ---- halo-reuse 3x3 kernel (csrc/tc_conv3x3.cuh) ---------------------------------------------------------------------
def run_conv_halo(x_nhwc, w_oihw, bias, relu, want_full=True, want_pool=False, out=None, out_ch_off=0):
    lib, check = _lib()
    b, h, w, c = x_nhwc.shape
    o, i, _, _ = w_oihw.shape
    wp = pack_conv_weight(w_oihw, i if i % 32 == 0 else (i + 63) // 64 * 64)   # C_in = 32 / 96: the SWIZZLE_64B (K = 32 blocks) variant
    full = pool = None
    if want_full:
        full = out if out is not None else torch.zeros(b, h, w, o, dtype=torch.float16, device="cuda")
    if want_pool:
        pool = torch.zeros(b, h // 2, w // 2, o, dtype=torch.float16, device="cuda")
    fptr = (full.data_ptr() + out_ch_off * 2) if full is not None else None
    check(lib.airfe_op_conv3x3(x_nhwc.data_ptr(), i, w, h, b, x_nhwc.stride(2), wp.data_ptr(), bias.data_ptr(), o, i, int(relu),
                               fptr, full.shape[-1] if full is not None else 0, pool.data_ptr() if pool is not None else None, o, None))
    torch.cuda.synchronize()
    return full, pool


@pytest.mark.parametrize("cfg", [
    # b, h, w, cin, cout
    (2, 64, 64, 64, 64),       # weights resident in smem, 2 strips
    (1, 48, 40, 64, 32),       # N = 32, ragged tiles (h, w not multiples of 16)
    (2, 32, 32, 32, 32),       # C_in = 32 (channel OOB fill)
    (1, 64, 64, 96, 128),      # streamed weights, C_in = 96 = 3 K blocks of 32 (SWIZZLE_64B, no K padding)
    (1, 32, 32, 72, 128),      # streamed weights, C_in = 72 -> padded to 128: 2 K blocks of 64
    (2, 32, 32, 128, 128),     # streamed weights, 2 strips
    (1, 32, 32, 256, 320),     # N = 320 -> two N tiles of 160, single strip
    (1, 16, 16, 128, 256),     # N = 256, single strip
    (3, 8, 8, 128, 128),       # 8x8 maps: half-empty tile
    (2, 16, 16, 128, 64),
])
def test_conv3x3_halo_matches_torch(cfg):
    b, h, w, cin, cout = cfg
    x, wt, bias = _mk(b, h, w, cin, cout, 3, seed=(hash(cfg) & 0xFFFF) + 7)
    full, pool = run_conv_halo(x, wt, bias, True, want_full=True, want_pool=True)
    ref = ref_conv(x, wt, bias, True).half()
    _close(full, ref)
    ref_pool = F.max_pool2d(full.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert torch.equal(pool.float(), ref_pool)     # pooling the fp16-rounded values is exact


@pytest.mark.parametrize("cfg", [(2, 64, 64, 64, 64), (1, 48, 40, 64, 32), (2, 32, 32, 32, 32), (1, 40, 72, 32, 64), (2, 40, 44, 128, 64), (1, 64, 64, 128, 64)])
def test_conv3x3_kx_fold_all_instantiations(cfg, monkeypatch):
    """AIRFE_CONV_FOLD=2 routes every eligible layer (C_in <= 64 with C_out = 32 / 64; C_in = 128 -> 64 as two resident N tiles of 32 with
    two K blocks) through tc_conv3x3_fold_kernel -- by default the C_out = 32 layers and the large 128 -> 64 maps use it (csrc/tc_conv3x3.cu).
    Full store, fused 2x2 max-pool and ragged tiles (width not a multiple of 14, height not of 8 / 16)."""
    monkeypatch.setenv("AIRFE_CONV_FOLD", "2")
    b, h, w, cin, cout = cfg
    x, wt, bias = _mk(b, h, w, cin, cout, 3, seed=(hash(cfg) & 0xFFFF) + 11)
    full, pool = run_conv_halo(x, wt, bias, True, want_full=True, want_pool=True)
    _close(full, ref_conv(x, wt, bias, True).half())
    ref_pool = F.max_pool2d(full.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert torch.equal(pool.float(), ref_pool)
    monkeypatch.setenv("AIRFE_CONV_FOLD", "0")
    full0, pool0 = run_conv_halo(x, wt, bias, True, want_full=True, want_pool=True)
    _close(full0, full)       # nine-tap kernel vs folded kernel: same values up to the fp32 summation order


def test_conv3x3_halo_pool_only_and_channel_offset():
    x, wt, bias = _mk(1, 64, 64, 64, 32, 3, seed=21)
    _, pool = run_conv_halo(x, wt, bias, True, want_full=False, want_pool=True)
    ref = F.max_pool2d(ref_conv(x, wt, bias, True).half().float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    _close(pool, ref.half())
    buf = torch.zeros(1, 64, 64, 96, dtype=torch.float16, device="cuda")
    run_conv_halo(x, wt, bias, False, want_full=True, out=buf, out_ch_off=32)
    _close(buf[..., 32:64], ref_conv(x, wt, bias, False).half())
    assert float(buf[..., :32].abs().max()) == 0 and float(buf[..., 64:].abs().max()) == 0
