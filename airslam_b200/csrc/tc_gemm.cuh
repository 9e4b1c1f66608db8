// tcgen05 implicit-GEMM convolution / batched GEMM for sm_100a.
//
// One kernel covers every dense contraction on the hot path (SURVEY.md §7.2 K2, K3, K11-MLP, K14 products, K15, K17):
//   D[pixel, n] = sum_{tap, c} A[pixel + offset(tap), c] * Bw[n, tap*C_pad + c]   (+ bias, ReLU)  -> fp16 / fp32 store
//  * A is an NHWC fp16 tensor seen through a 4-D TMA map (C, W, H, B); a tile of 128 output pixels is a
//    (tw x th x tb) box, and a 3x3 tap is the same box shifted by (dx,dy): TMA's out-of-bounds zero fill IS the
//    convolution padding.  The 128B-swizzled box lands in shared memory exactly in the K-major SWIZZLE_128B
//    canonical layout tcgen05.mma wants (128 rows x 64 fp16), so no repacking is ever done by threads.
//  * Bw is [N, K] K-major (weights packed tap-major) or, for attention P.V, an MN-major [K, N=64] tile.
//  * Accumulators live in TMEM (double-buffered: the epilogue of tile i overlaps the MMAs of tile i+1).
//  * Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+ TMEM owner), warps 2-5 = epilogue.
//  * Persistent: grid = min(#tiles, #SMs); tiles are strided over CTAs.
//  * BRES (weights-resident) mode: when the [block_n x K] weight panel fits, every CTA loads the panel of its own N tile once
//    (the grid is a multiple of n_tiles, so a CTA's N tile never changes) and the ring carries A tiles only.  Without it all
//    148 CTAs re-stream the same few KiB of weights from L2 for every 128-row tile, which bounded the 1x1 convolutions and the
//    matcher projections at 2-3 TB/s of L2 traffic (tools/prof_gemm.py).
//  * MN_MAJOR / BRES are template parameters so the single-warp MMA issue loop carries no runtime branches (see tc_conv3x3.cuh).
#pragma once
#include "ptx.cuh"
#include "tc_gemm_params.h"

namespace airfe {

// EW = epilogue warps per TMEM lane quarter (2 or 4): the epilogue is latency-bound with two warps per scheduler (ncu of the Wqkv GEMM: 68 %
// of the cycles no warp eligible; long-scoreboard waits on the rotary table and fixed-latency waits) -- EW = 4 gives each scheduler four
// epilogue warps.  Per-element arithmetic and therefore results are identical; only the assignment of 16-column chunks to warps changes.
// The TMA-store epilogue (off by default) exists for EW = 2 only.
template <bool MN_MAJOR, bool BRES, int EW = 2>
__global__ void __launch_bounds__(64 + 128 * EW, 1) tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A | B)] | barriers | tmem slot
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = tc_b_bytes(p.block_n, MN_MAJOR);
  const int stage_bytes = BRES ? kABytes : kABytes + b_bytes;
  const int ksteps = p.taps * p.kblocks;
  uint8_t* smem_bres = smem + p.stages * stage_bytes;                       // BRES: ksteps panels of b_bytes behind the A ring
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_bres + (BRES ? ksteps * b_bytes : 0));
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* bres_bar = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  uint8_t* smem_store = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 1023) & ~uintptr_t(1023));   // tma_store: 8 staging tiles of 4 KiB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_x = p.tiles_x;
  const int m_tiles = tiles_x * p.tiles_y * p.tiles_b;
  const int total_tiles = m_tiles * p.n_tiles;
  const uint32_t tmem_cols = tc_tmem_cols(p.block_n);
  const uint32_t acc_stride = tc_acc_stride(p.block_n);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmB);
    if (p.tma_store) ptx::prefetch_tmap(&p.tmO);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 4 * EW);
    }
    ptx::mbar_init(bres_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  __shared__ __align__(16) float s_bias[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_bias[i] = (p.bias && i < p.n_valid) ? p.bias[i] : 0.f;   // weights: not produced by a kernel
  __syncthreads();
  ptx::pdl_launch_dependents();   // programmatic dependent launch: see launch_pdl (common.h)
  ptx::pdl_wait();                // activations, device-side counts and residuals are read only below this line

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t b_tx = MN_MAJOR ? 64 * 128 : p.block_n * 128;
      const uint32_t tx_bytes = BRES ? kABytes : kABytes + b_tx;
      if (BRES) {   // this CTA's N tile is fixed (grid % n_tiles == 0): load its weight panel once
        const int nt = blockIdx.x % p.n_tiles;
        ptx::mbar_arrive_expect_tx(bres_bar, (uint32_t)ksteps * b_tx);
        for (int tap = 0; tap < p.taps; ++tap)
          for (int kb = 0; kb < p.kblocks; ++kb) {
            const int koff = tap * p.c_in_pad + kb * kBlockK;
            uint8_t* sb = smem_bres + (tap * p.kblocks + kb) * b_bytes;
            if (MN_MAJOR) ptx::tma_load_4d(sb, &p.tmB, bres_bar, nt * p.block_n, koff, 0, 0);
            else ptx::tma_load_4d(sb, &p.tmB, bres_bar, koff, nt * p.block_n, 0, 0);
          }
      }
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % p.n_tiles;
        const int mt = t / p.n_tiles;
        const int tx = mt % tiles_x;
        const int ty = (mt / tiles_x) % p.tiles_y;
        const int tz = mt / (tiles_x * p.tiles_y);
        const int x0 = tx * p.tw, y0 = ty * p.th, b0 = tz * p.tb;
        if (p.dyn_w && x0 >= __ldg(p.dyn_w + b0 * p.dyn_w_stride)) continue;   // rows beyond the device-side count: all roles skip alike
        const int bb = p.b_batched ? (b0 ^ p.b_batch_xor) : 0;
        const int hb = p.b_batched ? y0 : 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * stage_bytes;
            uint8_t* sb = sa + kABytes;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            ptx::tma_load_4d(sa, &p.tmA, &full_bar[stage], kb * kBlockK, x0 + dx, y0 + dy, b0);
            if (!BRES) {
              const int koff = tap * p.c_in_pad + kb * kBlockK;
              if (MN_MAJOR) ptx::tma_load_4d(sb, &p.tmB, &full_bar[stage], nt * p.block_n, koff, hb, bb);
              else ptx::tma_load_4d(sb, &p.tmB, &full_bar[stage], koff, nt * p.block_n, hb, bb);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the (uniform) loop, one elected lane issues =====
    const uint32_t idesc = ptx::make_idesc_f16(kTileM, p.block_n, MN_MAJOR ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint64_t d_const = ptx::smem_desc_base_sw128(1024);
    const uint64_t db_mn_const = (static_cast<uint64_t>((8192 >> 4) & 0x3FFF) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                                 (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
    const uint64_t da_base = d_const + (ptx::smem_u32(smem) >> 4);
    const uint64_t db_base = (MN_MAJOR ? db_mn_const : d_const) + (ptx::smem_u32(BRES ? smem_bres : smem + kABytes) >> 4);
    constexpr uint32_t bstep = MN_MAJOR ? (2048 >> 4) : 2;
    const uint32_t stage16 = (uint32_t)stage_bytes >> 4, b16 = (uint32_t)b_bytes >> 4;
    if (BRES) { ptx::mbar_wait(bres_bar, 0); ptx::tc_fence_after(); }
    // Optional (p.prewait, off by default): poll the barriers of unit u + 1 (next K step, or next tile: accumulator free + first stage landed)
    // BEFORE unit u is issued; with >= 3 ring stages u + 1 never depends on u.  Tested in round 2 against the idea that post-issue polling
    // idles the tensor pipe: no gain (profiles/r02_prewait_ab.txt) -- these GEMMs wait for data, not for the issuing thread.
    const bool prewait = p.prewait && p.stages >= 3;
    auto tile_valid = [&](int t) {
      if (!p.dyn_w) return true;
      const int mt = t / p.n_tiles;
      return (mt % tiles_x) * p.tw < __ldg(p.dyn_w + (mt / (tiles_x * p.tiles_y)) * p.tb * p.dyn_w_stride);
    };
    int t = blockIdx.x;
    while (t < total_tiles && !tile_valid(t)) t += gridDim.x;
    if (t < total_tiles) { ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1); ptx::mbar_wait(&full_bar[stage], phase); }
    while (t < total_tiles) {
      int tn = t + gridDim.x;
      while (tn < total_tiles && !tile_valid(tn)) tn += gridDim.x;
      const uint32_t d_tmem = tmem_base + acc * acc_stride;
      const int nacc = acc ^ 1;
      const uint32_t nacc_phase = nacc == 0 ? acc_phase ^ 1 : acc_phase;
      for (int ks = 0; ks < ksteps; ++ks) {
        const bool last = ks == ksteps - 1;
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == p.stages) { nstage = 0; nphase ^= 1; }
        if (prewait) {
          if (!last) ptx::mbar_wait(&full_bar[nstage], nphase);
          else if (tn < total_tiles) { ptx::mbar_wait(&tmem_empty[nacc], nacc_phase ^ 1); ptx::mbar_wait(&full_bar[nstage], nphase); }
        }
        ptx::tc_fence_after();
        const uint64_t da = da_base + (uint64_t)((uint32_t)stage * stage16);
        const uint64_t db = db_base + (uint64_t)(BRES ? (uint32_t)ks * b16 : (uint32_t)stage * stage16);
        const uint32_t first = ks != 0;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) ptx::umma_f16(d_tmem, da + 2 * k, db + bstep * k, idesc, k ? 1u : first);
          ptx::umma_commit(&empty_bar[stage]);
          if (last) ptx::umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (!prewait) {
          if (!last) ptx::mbar_wait(&full_bar[nstage], nphase);
          else if (tn < total_tiles) { ptx::mbar_wait(&tmem_empty[nacc], nacc_phase ^ 1); ptx::mbar_wait(&full_bar[nstage], nphase); }
        }
        stage = nstage; phase = nphase;
      }
      acc = nacc; acc_phase = nacc_phase;
      t = tn;
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (+bias, ReLU) -> global =====
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    constexpr int NBUF = EW == 2 ? 2 : 1;                         // register sets for the chunk pipeline
    const int part = (warp - 2) >> 2;                             // 0 .. EW - 1: this warp's share of the tile's columns
    const int chunks = p.block_n / 16;
    const int c_begin = (chunks * part + EW - 1) / EW, c_end = (chunks * (part + 1) + EW - 1) / EW;
    const int row = quarter * 32 + lane;
    // tma_store: this warp's 32 rows x 64 columns go through its own 4 KiB staging tile (row = lane, 128 bytes, 16-byte pieces XOR-swizzled
    // like SWIZZLE_128B so that the 32 lanes' 16-byte writes spread over all banks) and leave with ONE cp.async.bulk.tensor store per 64
    // columns: full 128-byte lines instead of 32 scattered 32-byte segments per STG.256 (the LSU data pipe was 42 % busy with those).
    uint8_t* stg = smem_store + ((warp - 2) & 7) * 4096;          // (tma_store is planned for EW = 2 only)
    const int r0w = quarter * 32;                                       // first tile row of this warp
    const int bx = r0w % p.tw, by = (r0w / p.tw) % p.th, bb = r0w / (p.tw * p.th);
    bool store_pending = false;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int nt = t % p.n_tiles;
      const int mt = t / p.n_tiles;
      const int tx = mt % tiles_x;
      const int ty = (mt / tiles_x) % p.tiles_y;
      const int tz = mt / (tiles_x * p.tiles_y);
      const int x = tx * p.tw + row % p.tw;
      const int y = ty * p.th + (row / p.tw) % p.th;
      const int b = tz * p.tb + row / (p.tw * p.th);
      const int vW = p.dyn_w ? min(__ldg(p.dyn_w + tz * p.tb * p.dyn_w_stride), p.W) : p.W;
      if (p.dyn_w && tx * p.tw >= vW) continue;
      const bool valid = (x < vW) && (y < p.H) && (b < p.B);
      const long long off = (long long)b * p.out_sb + (long long)y * p.out_sy + (long long)x * p.out_sx;
      const int n0 = nt * p.block_n;
      // fused rotary embedding: the cos / sin values of this keypoint for chunk cc + 1 are fetched while chunk cc is processed, the first
      // chunk's before the accumulator wait (ncu source view of the Wqkv GEMM: 30 % of the warp samples sat on the first FMUL after these
      // loads when they were issued at the point of use)
      float4 rv[NBUF][4];
      const float* rot_row = p.rot ? p.rot + ((long long)b * p.W + x) * 64 : nullptr;
      auto rot_fetch = [&](float4 (&dst)[4], int nb) {
        if (rot_row && valid && nb < p.rot_cols) {
          const float4* r4 = reinterpret_cast<const float4*>(rot_row + (nb & 63));
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = __ldg(r4 + i);
        }
      };
      rot_fetch(rv[0], n0 + c_begin * 16);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * acc_stride + (uint32_t(quarter * 32) << 16);
      uint32_t rr[NBUF][16];
      // one 16-column chunk of this thread's row: accumulators r, rotary values rvu (fetched ahead), first column c of the tile
      auto do_chunk = [&](const uint32_t (&r)[16], const float4 (&rvu)[4], const int c) {
        float v[16];
        float bsm[16];
        {
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + ((n0 + c) & 1023));
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float4 bb = b4[i]; bsm[4 * i] = bb.x; bsm[4 * i + 1] = bb.y; bsm[4 * i + 2] = bb.z; bsm[4 * i + 3] = bb.w; }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float f = __uint_as_float(r[i]);
          f += bsm[i];
          if (n0 + c + i < p.scale_cols) f *= p.scale;
          if (p.relu) f = fmaxf(f, 0.f);
          v[i] = f;
        }
        const int nbase = n0 + c;
        if (p.rot && nbase < p.rot_cols && valid) {
          // fused rotary embedding: adjacent columns (2j, 2j+1) of a 64-wide head rotate by the keypoint's angle j;
          // rot[row][2j] = cos, rot[row][2j+1] = sin  (lg_prepare_kernel), rows are (b, x) of the H == 1 layout
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 cs = rvu[i];
            const float a0 = v[4 * i], a1 = v[4 * i + 1], a2 = v[4 * i + 2], a3 = v[4 * i + 3];
            v[4 * i] = a0 * cs.x - a1 * cs.y;     v[4 * i + 1] = a1 * cs.x + a0 * cs.y;
            v[4 * i + 2] = a2 * cs.z - a3 * cs.w; v[4 * i + 3] = a3 * cs.z + a2 * cs.w;
          }
        }
        // column -> address: plain, or split into sections that live in separate matrices
        const long long ncol = p.out_split ? (long long)(nbase / p.out_split) * p.out_split_stride + (nbase % p.out_split) : (long long)nbase;
        if (valid && nbase < p.n_valid) {
          if (p.resid) {
            const float* rs = p.resid + off + nbase;
            if (nbase + 16 <= p.n_valid) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 q = *reinterpret_cast<const float4*>(rs + i);
                v[i] += q.x; v[i + 1] += q.y; v[i + 2] += q.z; v[i + 3] += q.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (nbase + i < p.n_valid) v[i] += __ldg(rs + i);   // static indices: v[] must stay in registers
            }
          }
          if (p.out2) {
            __half* o2 = reinterpret_cast<__half*>(p.out2) + (long long)b * p.out2_sb + (long long)y * p.out2_sy + (long long)x * p.out2_sx + nbase;
            if (nbase + 16 <= p.n_valid) {
              uint32_t h[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __half2 h2 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                h[i] = *reinterpret_cast<uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(o2) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(o2 + 8) = make_uint4(h[4], h[5], h[6], h[7]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (nbase + i < p.n_valid) o2[i] = __float2half_rn(v[i]);
            }
          }
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + off + ncol;
            if (nbase + 16 <= p.n_valid) {
              if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {
                const float lo[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]}, hi[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
                ptx::st_global_256f(o, lo);
                ptx::st_global_256f(o + 8, hi);
              } else {
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (nbase + i < p.n_valid) o[i] = v[i];
            }
          } else if (!p.tma_store) {
            __half* o = reinterpret_cast<__half*>(p.out) + off + ncol;
            if (nbase + 16 <= p.n_valid) {
              uint32_t h[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __half2 h2 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                h[i] = *reinterpret_cast<uint32_t*>(&h2);
              }
              if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {
                ptx::st_global_256(o, h);
              } else {
                *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(o + 8) = make_uint4(h[4], h[5], h[6], h[7]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (nbase + i < p.n_valid) o[i] = __float2half_rn(v[i]);
            }
          }
        }
        if (p.tma_store) {
          // every lane stages its 16 values (rows beyond the valid extent as zeros: their positions inside the tensor may be read as padding
          // rows of a later tile, and must stay finite), the group of four chunks = 64 columns leaves as one box
          const int cg = ((c >> 4) - c_begin) & 3;
          if (cg == 0) {
            if (store_pending) { if (lane == 0) ptx::tma_store_wait_read(); store_pending = false; }
            __syncwarp();
          }
          uint32_t h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            __half2 h2 = valid ? __floats2half2_rn(v[2 * i], v[2 * i + 1]) : __floats2half2_rn(0.f, 0.f);
            h[i] = *reinterpret_cast<uint32_t*>(&h2);
          }
          uint8_t* srow = stg + lane * 128;
          *reinterpret_cast<uint4*>(srow + (((2 * cg) ^ (lane & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(srow + (((2 * cg + 1) ^ (lane & 7)) << 4)) = make_uint4(h[4], h[5], h[6], h[7]);
          if (cg == 3) {
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              const int n_first = nbase - 48;                                 // first column of the 64-column group
              const int sec = p.out_split ? n_first / p.out_split : 0;
              ptx::tma_store_5d(&p.tmO, stg, p.out_split ? n_first % p.out_split : n_first, tx * p.tw + bx, ty * p.th + by, tz * p.tb + bb, sec);
              ptx::tma_store_commit();
            }
            store_pending = true;
          }
        }
      };
      if constexpr (NBUF == 2) {
        // EW = 2: the TMEM load and the rotary values of chunk cc + 1 are in flight while chunk cc is processed
        if (c_begin < c_end) ptx::tmem_ld16(taddr + c_begin * 16, rr[0]);
#pragma unroll 1
        for (int cq = c_begin; cq < c_end; cq += 2)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (cq + u >= c_end) break;
            const int c = (cq + u) * 16;
            ptx::tmem_ld_wait();
            if (cq + u + 1 < c_end) { ptx::tmem_ld16(taddr + c + 16, rr[u ^ 1]); rot_fetch(rv[u ^ 1], n0 + c + 16); }
            do_chunk(rr[u], rv[u], c);
          }
      } else {
        // EW = 4: single-buffered (112 registers per thread at 576 threads); four warps per scheduler hide the load latencies instead
#pragma unroll 1
        for (int cc = c_begin; cc < c_end; ++cc) {
          const int c = cc * 16;
          if (cc != c_begin) rot_fetch(rv[0], n0 + c);
          ptx::tmem_ld16(taddr + c, rr[0]);
          ptx::tmem_ld_wait();
          do_chunk(rr[0], rv[0], c);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  if (p.tma_store && warp >= 2 && lane == 0) ptx::tma_store_wait_all();      // the staging tiles must outlive the bulk stores that read them
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace airfe
