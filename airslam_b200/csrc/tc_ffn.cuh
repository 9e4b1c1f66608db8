// Fused LightGlue transformer-block tail on tcgen05 (G4: out_proj / to_out -> ffn.0 -> LayerNorm -> GELU -> ffn.3 -> residual).
// Per 128-keypoint tile a CTA chains three GEMMs whose intermediates never leave the SM:
//   phase 1  msg = ctx . Wout^T + b           (K = 256, N = 256)   TMEM -> fp16 -> shared memory (A operand, K-major SWIZZLE_128B)
//   phase 2  h   = [x | msg] . W0^T + b0      (K = 512, N = 512)   accumulators fill all 512 TMEM columns
//            LayerNorm(512, eps 1e-5) + exact GELU over TMEM rows -> fp16 -> shared memory (overwrites [x | msg])
//   phase 3  y   = gelu . W3^T + b3           (K = 512, N = 256)   x += y  (fp32 residual stream + fp16 operand copy to HBM)
// Weights (896 KB per tile) stream from L2 through a 5-stage TMA ring of [128 x 64] tiles.  The unfused path needed four GEMM
// launches + one LayerNorm launch per block and moved msg / h / gelu(h) through HBM.
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2-9: epilogues / LayerNorm: two warps per TMEM lane quarter, each thread owns one
//   keypoint row and one half of the columns; the LayerNorm statistics of the two halves meet in shared memory (named barrier 1)
#pragma once
#include "ptx.cuh"

namespace airfe {

struct FfnParams {
  CUtensorMap tmCtx;   // 4-D (256, cap, 1, slots)  box (64, 128, 1, 1)   attention context (fp16)
  CUtensorMap tmX16;   // 4-D (256, cap, 1, slots)  box (64, 128, 1, 1)   fp16 operand copy of x (row stride 512)
  CUtensorMap tmWo;    // 4-D (256, 256, 1, 1)      box (64, 128, 1, 1)
  CUtensorMap tmW0;    // 4-D (512, 512, 1, 1)
  CUtensorMap tmW3;    // 4-D (512, 256, 1, 1)
  const float *b_out, *b0, *b3, *ln_g, *ln_b;
  float* x;            // [slots*cap][256] fp32 residual stream (in/out)
  __half* x16;         // fp16 copy (row stride 512 elements)
  const int* n;        // [slots]
  int slots, cap;
  long long* trace;    // authoring aid (airfe_debug_match_trace): CTA 0 writes clock64 stamps of its first 8 row tiles, 16 slots per tile
  int prewait;         // MMA issuer polls the next weight tile's barrier before issuing the current one (0 with AIRFE_PREWAIT=1 switches it on)
};

constexpr int kFfnThreads = 320;
constexpr int kFfnWideThreads = 576;   // EW = 4: sixteen epilogue warps
constexpr int kFfnBStages = 5;      // 16 KiB weight tiles in flight per SM (226.5 KiB of shared memory in total)
constexpr int kFfnSmemBytes = 8 * 16384 + kFfnBStages * 16384 + (2048 + 1024) * 4 + 1024 + 256;

// RELU = false: LightGlue (LayerNorm + exact GELU between the two FFN GEMMs).  RELU = true: SuperGlue's MLP([x | merge(ctx)]) = 512 -> 512
// (BatchNorm folded into the weights) -> ReLU -> 256 with the same residual: identical GEMM shapes, a one-pass phase-2 epilogue.
// EW = epilogue warps per TMEM lane quarter (2 or 4); a thread owns one keypoint row and 1 / EW of the columns of every phase.  The three GEMM
// phases and their epilogues are serial per row tile (phase 2 fills all 512 TMEM columns), so during the epilogues the SM runs only these
// warps: ncu of the EW = 2 kernel shows 2.5 warps per scheduler, 74 % of the cycles without an eligible warp and 26 % of the issue slots
// used -- EW = 4 halves the work per thread and doubles the warps that hide each other's TMEM / global / erff latencies.
template <bool RELU, int EW = 2>
__global__ void __launch_bounds__(64 + 128 * EW, 1) tc_ffn_kernel(const __grid_constant__ FfnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // 8 K blocks of [128 rows x 64] fp16
  uint8_t* sB = sA + 8 * 16384;                         // weight ring
  float* sPar = reinterpret_cast<float*>(sB + kFfnBStages * 16384);   // b_out[256] b0[512] b3[256] g[512] beta[512]
  float* sRed = sPar + 2048;                             // LayerNorm partials: sum[EW][128], var[EW][128] (var at + 512)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 1024);
  uint64_t* b_full = bars;                              // [4]
  uint64_t* b_empty = bars + kFfnBStages;               // [4]
  uint64_t* ctx_full = bars + 2 * kFfnBStages;
  uint64_t* x_full = ctx_full + 1;
  uint64_t* a_free = ctx_full + 2;                      // all MMAs of the tile retired: A blocks may be reloaded
  uint64_t* acc_full = ctx_full + 3;                    // a GEMM phase finished: accumulators valid
  uint64_t* epi_done = ctx_full + 4;                    // (count 4 * EW) epilogue finished with TMEM and with its smem writes
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctx_full + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_slot = p.cap >> 7;
  // Weight tiles are consumed in the same (n tile, k block) order by every CTA: a row tile's K accumulation order must not depend on which
  // CTA gets it, so that a pair inside a batch of 32 is bit-identical to the same pair alone (tests/test_batch_invariance_gpu.py).  A per-CTA
  // rotation of that order was measured in round 1 (L2 hot-spot theory) and bought nothing.
  const int total_tiles = p.slots * tiles_per_slot;

  for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
    float v;
    if (i < 256) v = p.b_out[i];
    else if (i < 768) v = p.b0[i - 256];
    else if (i < 1024) v = p.b3[i - 768];
    else if (i < 1536) v = p.ln_g[i - 1024];
    else v = p.ln_b[i - 1536];
    sPar[i] = v;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmCtx); ptx::prefetch_tmap(&p.tmX16); ptx::prefetch_tmap(&p.tmWo); ptx::prefetch_tmap(&p.tmW0); ptx::prefetch_tmap(&p.tmW3);
    for (int i = 0; i < kFfnBStages; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    ptx::mbar_init(ctx_full, 1); ptx::mbar_init(x_full, 1); ptx::mbar_init(a_free, 1); ptx::mbar_init(acc_full, 1); ptx::mbar_init(epi_done, 4 * EW);
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_launch_dependents();   // programmatic dependent launch: see launch_pdl (common.h); only weights / biases were read so far
  ptx::pdl_wait();
  const float* s_bout = sPar; const float* s_b0 = sPar + 256; const float* s_b3 = sPar + 768; const float* s_g = sPar + 1024; const float* s_be = sPar + 1536;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int sb = 0; uint32_t pb = 0, pt = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int slot = t / tiles_per_slot, r0 = (t % tiles_per_slot) * 128;
        if (r0 >= __ldg(p.n + slot)) continue;
        ptx::mbar_wait(a_free, pt ^ 1);
        ptx::mbar_arrive_expect_tx(ctx_full, 4u * 16384u);
        for (int kb = 0; kb < 4; ++kb) ptx::tma_load_4d(sA + (4 + kb) * 16384, &p.tmCtx, ctx_full, kb * 64, r0, 0, slot);
        ptx::mbar_arrive_expect_tx(x_full, 4u * 16384u);
        for (int kb = 0; kb < 4; ++kb) ptx::tma_load_4d(sA + kb * 16384, &p.tmX16, x_full, kb * 64, r0, 0, slot);
        pt ^= 1;
        // weight tiles in the order the MMA warp consumes them: phase 1 (Wout: 2 n x 4 k), phase 2 (W0: 4 n x 8 k), phase 3 (W3: 2 n x 8 k)
        for (int ph = 0; ph < 3; ++ph) {
          const CUtensorMap* tm = ph == 0 ? &p.tmWo : (ph == 1 ? &p.tmW0 : &p.tmW3);
          const int nn = ph == 1 ? 4 : 2, nk = ph == 0 ? 4 : 8;
          for (int nt = 0; nt < nn; ++nt)
            for (int kb = 0; kb < nk; ++kb) {
              const int ntr = nt, kbr = kb;
              ptx::mbar_wait(&b_empty[sb], pb ^ 1);
              ptx::mbar_arrive_expect_tx(&b_full[sb], 16384u);
              ptx::tma_load_4d(sB + sb * 16384, tm, &b_full[sb], kbr * 64, ntr * 128, 0, 0);
              if (++sb == kFfnBStages) { sb = 0; pb ^= 1; }
            }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint64_t d_const = ptx::smem_desc_base_sw128(1024);
    const uint32_t idesc = ptx::make_idesc_f16(128, 128, 0);
    const uint64_t da0 = d_const + (ptx::smem_u32(sA) >> 4);
    int sb = 0; uint32_t pb = 0, pt = 0, pepi = 0;
    int tcount = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int slot = t / tiles_per_slot, r0 = (t % tiles_per_slot) * 128;
      if (r0 >= __ldg(p.n + slot)) continue;
      long long* trp = (p.trace && blockIdx.x == 0 && lane == 0 && tcount < 8) ? p.trace + tcount * 16 : nullptr;
      ++tcount;
      for (int ph = 0; ph < 3; ++ph) {
        // operands ready?  phase 1: ctx tile landed and the previous tile's last epilogue released TMEM; phase 2: msg written + x tile
        // landed; phase 3: gelu(h) written.  epi_done completes once per phase (three times per tile).
        if (ph == 0) { ptx::mbar_wait(ctx_full, pt); ptx::mbar_wait(epi_done, pepi ^ 1); }
        else { ptx::mbar_wait(epi_done, pepi ^ 1); if (ph == 1) ptx::mbar_wait(x_full, pt); }
        pepi ^= 1;
        ptx::tc_fence_after();
        if (trp) trp[2 * ph] = clock64();             // operands of this phase ready
        const int nn = ph == 1 ? 4 : 2, nk = ph == 0 ? 4 : 8, a_first = ph == 0 ? 4 : 0;
        // weight tiles: the barrier of the NEXT tile of the ring is polled before the MMAs of the current one are issued (the issuing thread is
        // back-pressured, so anything it does between two issue blocks is idle time of the tensor pipe; 4 ring stages: tile u + 1 never
        // depends on tile u).  The first tile of a row tile is waited for at the phase-1 start.
        if (ph == 0) ptx::mbar_wait(&b_full[sb], pb);
        for (int nt = 0; nt < nn; ++nt)
          for (int kb = 0; kb < nk; ++kb) {
            int nsb = sb + 1;
            uint32_t npb = pb;
            if (nsb == kFfnBStages) { nsb = 0; npb ^= 1; }
            const bool more = !(ph == 2 && nt == nn - 1 && kb == nk - 1);        // another weight tile of THIS row tile follows
            if (p.prewait && more) ptx::mbar_wait(&b_full[nsb], npb);
            ptx::tc_fence_after();
            const int ntr = nt, kbr = kb;
            if (ptx::elect_one()) {
              const uint64_t da = da0 + (uint64_t)((a_first + kbr) * (16384 >> 4));
              const uint64_t db = d_const + (ptx::smem_u32(sB + sb * 16384) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base + ntr * 128, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
              ptx::umma_commit(&b_empty[sb]);
            }
            __syncwarp();
            if (!p.prewait && more) ptx::mbar_wait(&b_full[nsb], npb);
            sb = nsb; pb = npb;
          }
        if (ptx::elect_one()) {
          ptx::umma_commit(acc_full);
          if (ph == 2) ptx::umma_commit(a_free);
        }
        __syncwarp();
        if (trp) trp[2 * ph + 1] = clock64();         // all MMAs of this phase issued
      }
      pt ^= 1;
    }
  } else {
    // ===== epilogues: thread = one keypoint row x one part (1 / EW) of the columns =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;          // column part 0 .. EW - 1 (EW = 2: the low / high half)
    constexpr int P1 = 256 / EW, P2 = 512 / EW;  // columns per thread: phases 1 and 3 (N = 256), phase 2 (N = 512)
    const int row = quarter * 32 + lane;
    const uint32_t trow = tmem_base + (uint32_t(quarter * 32) << 16);
    uint32_t pacc = 0;
    int tcount = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int slot = t / tiles_per_slot, r0 = (t % tiles_per_slot) * 128;
      const int ns = __ldg(p.n + slot);
      if (r0 >= ns) continue;
      long long* tre = (p.trace && blockIdx.x == 0 && warp == 2 && lane == 0 && tcount < 8) ? p.trace + tcount * 16 + 6 : nullptr;
      ++tcount;
      const bool valid = (r0 + row) < ns;
      const long long grow = (long long)slot * p.cap + r0 + row;
      // ---- phase 1 epilogue: msg = acc + b_out -> fp16 -> A blocks 4..7 (this warp: columns half*P1 .. +P1) ----
      ptx::mbar_wait(acc_full, pacc); pacc ^= 1;
      ptx::tc_fence_after();
      if (tre) tre[0] = clock64();                    // phase-1 accumulators complete
#pragma unroll 1
      for (int c = half * P1; c < half * P1 + P1; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(trow + c, r);
        ptx::tmem_ld_wait();
        uint8_t* dst = sA + (4 + (c >> 6)) * 16384 + row * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = ch * 8 + e * 2;
            __half2 h2 = __floats2half2_rn(__uint_as_float(r[j]) + s_bout[c + j], __uint_as_float(r[j + 1]) + s_bout[c + j + 1]);
            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
          const int chunk = ((c & 63) >> 3) + ch;
          *reinterpret_cast<uint4*>(dst + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(epi_done);
      if (tre) tre[1] = clock64();                    // phase-1 epilogue done (msg in shared memory)
      // ---- phase 2 epilogue: LayerNorm + GELU over 512 columns (this warp: columns half*256 .. +256) -> fp16 -> A blocks 0..7 ----
      ptx::mbar_wait(acc_full, pacc); pacc ^= 1;
      ptx::tc_fence_after();
      if (tre) tre[2] = clock64();                    // phase-2 accumulators complete
      const int c_lo = half * P2, c_hi = c_lo + P2;
      if constexpr (RELU) {
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) {
          uint32_t r[32];
          ptx::tmem_ld32(trow + c, r);
          ptx::tmem_ld_wait();
          uint8_t* dst = sA + (c >> 6) * 16384 + row * 128;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = ch * 8 + e * 2;
              __half2 h2 = __floats2half2_rn(fmaxf(__uint_as_float(r[j]) + s_b0[c + j], 0.f), fmaxf(__uint_as_float(r[j + 1]) + s_b0[c + j + 1], 0.f));
              pk[e] = *reinterpret_cast<uint32_t*>(&h2);
            }
            const int chunk = ((c & 63) >> 3) + ch;
            *reinterpret_cast<uint4*>(dst + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      } else {
      // Three passes over this thread's 256 accumulator columns (mean, variance, normalise + GELU).  The TMEM loads are double-buffered:
      // the load of block i + 1 is in flight while block i is reduced.  The loops are unrolled by two only (static register-set indices):
      // fully unrolled, the kernel grew to 170 KB of SASS and ran 30 % SLOWER (instruction-cache misses, profiles/r02c_match_trace.txt).
      uint32_t rr[EW == 2 ? 2 : 1][32];
      float sum = 0.f;
      auto sum_blk = [&](const uint32_t (&r)[32], int c) {
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) part += __uint_as_float(r[j]) + s_b0[c + j];
        sum += part;
      };
      if constexpr (EW == 2) {
        ptx::tmem_ld32(trow + c_lo, rr[0]);
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 64) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld32(trow + c + 32, rr[1]);
          sum_blk(rr[0], c);
          ptx::tmem_ld_wait();
          ptx::tmem_ld32(trow + (c + 64 < c_hi ? c + 64 : c_lo), rr[0]);         // last iteration: first block of the next pass
          sum_blk(rr[1], c + 32);
        }
      } else {      // EW = 4: one register set (96 registers per thread at 576 threads); the other three warps of the scheduler cover the TMEM load
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) { ptx::tmem_ld32(trow + c, rr[0]); ptx::tmem_ld_wait(); sum_blk(rr[0], c); }
      }
      sRed[half * 128 + row] = sum;
      asm volatile("bar.sync 1, %0;" ::"n"(128 * EW) : "memory");
      if (tre) tre[3] = clock64();                    // LayerNorm pass 1 (mean) done
      // partial sums of the column parts, added in the same fixed order by every thread of the row
      const float mean = (EW == 2 ? sRed[row] + sRed[128 + row] : (sRed[row] + sRed[128 + row]) + (sRed[256 + row] + sRed[384 + row])) * (1.f / 512.f);
      float var = 0.f;
      auto var_blk = [&](const uint32_t (&r)[32], int c) {
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = __uint_as_float(r[j]) + s_b0[c + j] - mean; part = fmaf(d, d, part); }
        var += part;
      };
      if constexpr (EW == 2) {
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 64) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld32(trow + c + 32, rr[1]);
          var_blk(rr[0], c);
          ptx::tmem_ld_wait();
          ptx::tmem_ld32(trow + (c + 64 < c_hi ? c + 64 : c_lo), rr[0]);
          var_blk(rr[1], c + 32);
        }
      } else {
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) { ptx::tmem_ld32(trow + c, rr[0]); ptx::tmem_ld_wait(); var_blk(rr[0], c); }
      }
      sRed[512 + half * 128 + row] = var;
      asm volatile("bar.sync 1, %0;" ::"n"(128 * EW) : "memory");
      if (tre) tre[4] = clock64();                    // pass 2 (variance) done
      const float vsum = EW == 2 ? sRed[512 + row] + sRed[640 + row] : (sRed[512 + row] + sRed[640 + row]) + (sRed[768 + row] + sRed[896 + row]);
      const float rstd = 1.f / sqrtf(vsum * (1.f / 512.f) + 1e-5f);
      auto gelu_blk = [&](const uint32_t (&r)[32], int c) {
        uint8_t* dst = sA + (c >> 6) * 16384 + row * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float g2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int j = ch * 8 + e * 2 + u;
              const float y = (__uint_as_float(r[j]) + s_b0[c + j] - mean) * rstd * s_g[c + j] + s_be[c + j];
              g2[u] = 0.5f * y * (1.f + erff(y * 0.70710678118654752440f));
            }
            __half2 h2 = __floats2half2_rn(g2[0], g2[1]);
            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
          const int chunk = ((c & 63) >> 3) + ch;
          *reinterpret_cast<uint4*>(dst + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      };
      if constexpr (EW == 2) {
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 64) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld32(trow + c + 32, rr[1]);
          gelu_blk(rr[0], c);
          ptx::tmem_ld_wait();
          if (c + 64 < c_hi) ptx::tmem_ld32(trow + c + 64, rr[0]);
          gelu_blk(rr[1], c + 32);
        }
      } else {
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) { ptx::tmem_ld32(trow + c, rr[0]); ptx::tmem_ld_wait(); gelu_blk(rr[0], c); }
      }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(epi_done);
      if (tre) tre[5] = clock64();                    // phase-2 epilogue done (gelu(h) in shared memory)
      // ---- phase 3 epilogue: x += acc + b3 ; fp32 residual + fp16 operand copy (this warp: columns half*P1 .. +P1) ----
      ptx::mbar_wait(acc_full, pacc); pacc ^= 1;
      ptx::tc_fence_after();
      if (tre) tre[6] = clock64();                    // phase-3 accumulators complete
      // The residual rows come from L2 (written by the previous block's launch).  All four 32-byte loads of a 32-column group are issued
      // together and the next group's loads fly while this one is summed and stored: the first version issued them one at a time between
      // dependent stores -- sixteen serialised L2 round trips, 27.8 k cycles per tile (profiles/r02c_match_trace.txt).
      float* xr = p.x + grow * 256;
      __half* x16r = p.x16 + grow * 512;
      const int c3 = half * P1;
      float xv[EW == 2 ? 2 : 1][4][8];
      uint32_t r3[EW == 2 ? 2 : 1][32];
      auto ld_x = [&](float (&dstv)[4][8], int c) {
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) ptx::ld_global_256f_nv(xr + c + 8 * j, dstv[j]);      // rows are 1 KiB, columns multiples of 8 floats: 32-byte aligned
        }
      };
      auto add_store = [&](const float (&xin)[4][8], const uint32_t (&r)[32], int c) {
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = xin[j][e] + (__uint_as_float(r[8 * j + e]) + s_b3[c + 8 * j + e]);
            ptx::st_global_256f(xr + c + 8 * j, v);
            uint32_t h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { __half2 h2 = __floats2half2_rn(v[2 * e], v[2 * e + 1]); h[e] = *reinterpret_cast<uint32_t*>(&h2); }
            *reinterpret_cast<uint4*>(x16r + c + 8 * j) = make_uint4(h[0], h[1], h[2], h[3]);
          }
        }
      };
      if constexpr (EW == 2) {
        ld_x(xv[0], c3);
        ptx::tmem_ld32(trow + c3, r3[0]);
#pragma unroll 1
        for (int c = c3; c < c3 + P1; c += 64) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld32(trow + c + 32, r3[1]);
          ld_x(xv[1], c + 32);
          add_store(xv[0], r3[0], c);
          ptx::tmem_ld_wait();
          if (c + 64 < c3 + P1) { ptx::tmem_ld32(trow + c + 64, r3[0]); ld_x(xv[0], c + 64); }
          add_store(xv[1], r3[1], c + 32);
        }
      } else {
#pragma unroll 1
        for (int c = c3; c < c3 + P1; c += 32) {
          ld_x(xv[0], c);
          ptx::tmem_ld32(trow + c, r3[0]);
          ptx::tmem_ld_wait();
          add_store(xv[0], r3[0], c);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(epi_done);
      if (tre) tre[7] = clock64();                    // phase-3 epilogue done (x stored)
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

}  // namespace airfe
