// Fused multi-head attention for the matchers on tcgen05: softmax(Q K^T * scale) V for one (image slot, head)
// per CTA, 4 heads x 64 dims, up to 512 keys.  The score tile S (128 queries x n keys, fp32) lives ONLY in TMEM, the
// probabilities P only in shared memory (fp16, written by the softmax warps directly in the K-major SWIZZLE_128B layout the
// MMA reads), so nothing but Q, K, V and the context ever touches HBM -- the unfused path moved 67 MB of fp32 scores per
// attention through HBM three times.
//   warp 0 : TMA producer (K, V of the attended slot once per CTA; one Q tile per 128 queries)
//   warp 1 : MMA issuer   (S = Q K^T: 2 x N<=256 ; O = P V: per 64-key block, V as MN-major B operand)
//   warps 2-9 : softmax + epilogue: two warps per TMEM lane quarter; a thread owns one query row and every second 64-key block
//               (row max and row sum of the two halves meet in shared memory, named barrier 1), then half of the O columns -> fp16 ctx
// Arithmetic: fp16 operands, fp32 accumulate, fp32 softmax statistics; the tensor core sees fp16(exp(s - max)) and the 1/sum
// normalisation is applied to the fp32 output (same relative rounding as rounding the normalised probabilities).
#pragma once
#include "ptx.cuh"

namespace airfe {

struct AttnParams {
  CUtensorMap tmQ;   // 4-D (64, rows=cap, 4 heads, slots)  box (64, 128, 1, 1)
  CUtensorMap tmK;   // same geometry on the key matrix,     box (64, 256, 1, 1)
  CUtensorMap tmV;   // same geometry on the value matrix,   box (64, 256, 1, 1)   (rows = keys: MN-major B for P.V)
  const int* n;      // [slots] keypoints per slot (device)
  int slots, cap, slot_xor;
  const int* row_off; // optional packed layout: slot s owns rows [row_off[s], +n[s]) of the q / k / v / ctx matrices (tensor maps then have ONE slot of slots*cap rows)
  int q_split;       // CTAs per (slot, head): CTA part handles query tiles part, part + q_split, ... (small batches: more CTAs than slots x 4)
  float scale;       // applied to S before the softmax (1 for LightGlue: q,k pre-scaled; 1/8 for SuperGlue)
  __half* ctx;       // [slots*cap][256] fp16, head h at columns h*64
  long long* trace;  // authoring aid (airfe_debug_match_trace): CTA 0 writes clock64 stamps of its query tiles, 16 slots per tile
};

constexpr int kAttnThreads = 320;               // NP = 2: warps 2-9 = eight softmax warps
constexpr int kAttnThreadsWide = 576;           // NP = 4: warps 2-17 = sixteen softmax warps (four per scheduler)
constexpr int kAttnSmemQ = 128 * 128;            // 16 KiB
constexpr int kAttnSmemKV = 512 * 128;           // 64 KiB each
constexpr int kAttnPSlots = 4;                   // ring of 64-key P blocks (16 KiB each)
constexpr int kAttnSmemBytes = kAttnSmemQ + 2 * kAttnSmemKV + kAttnPSlots * 16384 + 4096 + 1024 + 256;

// NP = softmax warps per TMEM lane quarter (2 or 4): a thread owns one query row and every NP-th 64-key block.  The softmax is bound by
// MUFU.EX2 (16 / clk / SM) and by instruction latency, not by the tensor pipe (ncu: issue slot 34 % busy with 2.5 resident warps per
// scheduler), so NP = 4 doubles the warps that hide each other's latency; it runs single-buffered TMEM loads to stay under 113 registers.
template <int NP>
__global__ void __launch_bounds__(64 + NP * 128, 1) tc_attn_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnSmemQ;
  uint8_t* sV = sK + kAttnSmemKV;
  uint8_t* sP = sV + kAttnSmemKV;
  float* sRed = reinterpret_cast<float*>(sP + kAttnPSlots * 16384);   // row max [NP][128], row sum [NP][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 1024);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = bars + 2;
  uint64_t* s_full = bars + 3;
  uint64_t* o_full = bars + 4;
  uint64_t* s_free = bars + 5;
  uint64_t* p_full = bars + 6;                 // [4]
  uint64_t* p_empty = bars + 6 + kAttnPSlots;  // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * kAttnPSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int part = blockIdx.x % p.q_split, sh = blockIdx.x / p.q_split;
  const int slot = sh >> 2, head = sh & 3;
  const int kslot = slot ^ p.slot_xor;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmQ); ptx::prefetch_tmap(&p.tmK); ptx::prefetch_tmap(&p.tmV);
    ptx::mbar_init(kv_full, 1); ptx::mbar_init(q_full, 1); ptx::mbar_init(q_empty, 1);
    ptx::mbar_init(s_full, 1); ptx::mbar_init(o_full, 1); ptx::mbar_init(s_free, 4 * NP);
    for (int i = 0; i < kAttnPSlots; ++i) { ptx::mbar_init(&p_full[i], 4); ptx::mbar_init(&p_empty[i], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_launch_dependents();   // programmatic dependent launch: see launch_pdl (common.h)
  ptx::pdl_wait();                // the keypoint counts and Q / K / V are produced by earlier kernels: read them only from here on
  const int q_row0 = p.row_off ? __ldg(p.row_off + slot) : 0, k_row0 = p.row_off ? __ldg(p.row_off + kslot) : 0;   // packed: row bases inside the single slot
  const int q_slot = p.row_off ? 0 : slot, k_slot = p.row_off ? 0 : kslot;
  const int nq = min(__ldg(p.n + slot), p.cap);
  const int nk = min(__ldg(p.n + kslot), 512);
  const int q_tiles_all = (nq + 127) >> 7;
  const int q_tiles = q_tiles_all > part ? (q_tiles_all - part + p.q_split - 1) / p.q_split : 0;     // tiles of this CTA: t_global = part + t * q_split
  const int nkb = (nk + 63) >> 6;              // 64-key blocks of P / V
  const int nk16 = (nk + 15) & ~15;            // key extent of the S MMAs

  if (q_tiles > 0 && nk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        // ===== TMA producer =====
        const uint32_t kv_rows = (nk > 256) ? 512u : 256u;
        ptx::mbar_arrive_expect_tx(kv_full, 2u * kv_rows * 128u);
        ptx::tma_load_4d(sK, &p.tmK, kv_full, 0, k_row0, head, k_slot);      // packed: rows beyond nk belong to the next slot -- masked by the softmax
        ptx::tma_load_4d(sV, &p.tmV, kv_full, 0, k_row0, head, k_slot);
        if (nk > 256) {
          ptx::tma_load_4d(sK + 256 * 128, &p.tmK, kv_full, 0, k_row0 + 256, head, k_slot);
          ptx::tma_load_4d(sV + 256 * 128, &p.tmV, kv_full, 0, k_row0 + 256, head, k_slot);
        }
        uint32_t ph = 0;
        for (int t = 0; t < q_tiles; ++t) {
          ptx::mbar_wait(q_empty, ph ^ 1);
          ptx::mbar_arrive_expect_tx(q_full, 128u * 128u);
          ptx::tma_load_4d(sQ, &p.tmQ, q_full, 0, q_row0 + (part + t * p.q_split) * 128, head, q_slot);
          ph ^= 1;
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer (warp-uniform loop, elected lane issues) =====
      const uint64_t dk_const = ptx::smem_desc_base_sw128(1024);
      const uint64_t dv_const = (static_cast<uint64_t>((8192 >> 4) & 0x3FFF) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                                (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);       // MN-major SW128
      const uint64_t dq = dk_const + (ptx::smem_u32(sQ) >> 4);
      const uint64_t dk = dk_const + (ptx::smem_u32(sK) >> 4);
      const uint64_t dv = dv_const + (ptx::smem_u32(sV) >> 4);
      const int n_lo = nk16 < 256 ? nk16 : 256, n_hi = nk16 - n_lo;
      const uint32_t idesc_lo = ptx::make_idesc_f16(128, n_lo, 0);
      const uint32_t idesc_hi = ptx::make_idesc_f16(128, n_hi > 0 ? n_hi : 16, 0);
      const uint32_t idesc_pv = ptx::make_idesc_f16(128, 64, 1);
      ptx::mbar_wait(kv_full, 0);
      uint32_t ph = 0;
      uint32_t pph = 0;   // phase bit per P ring slot
      for (int t = 0; t < q_tiles; ++t) {
        long long* trp = (p.trace && blockIdx.x == 0 && lane == 0 && t < 8) ? p.trace + t * 16 : nullptr;
        if (trp) trp[0] = clock64();             // MMA warp reaches the tile (K / V landed)
        ptx::mbar_wait(q_full, ph);
        ptx::mbar_wait(s_free, ph ^ 1);          // TMEM free again (previous tile's O has been read)
        ptx::tc_fence_after();
        if (trp) trp[1] = clock64();             // Q landed, TMEM free
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base, dq + 2 * k, dk + 2 * k, idesc_lo, k != 0);
          if (n_hi > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base + 256, dq + 2 * k, dk + (256 * 128 >> 4) + 2 * k, idesc_hi, k != 0);
          }
          ptx::umma_commit(q_empty);             // Q tile consumed once these MMAs retire
          ptx::umma_commit(s_full);
        }
        __syncwarp();
        for (int kb = 0; kb < nkb; ++kb) {
          const int ps = kb & (kAttnPSlots - 1);
          ptx::mbar_wait(&p_full[ps], (pph >> ps) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint64_t dp = dk_const + (ptx::smem_u32(sP + ps * 16384) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16(tmem_base, dp + 2 * k, dv + (uint64_t)(((kb * 64 + k * 16) * 128) >> 4), idesc_pv, (kb | k) != 0);
            ptx::umma_commit(&p_empty[ps]);
            if (kb == nkb - 1) ptx::umma_commit(o_full);
          }
          __syncwarp();
          pph ^= 1u << ps;
        }
        if (trp) trp[2] = clock64();             // last P.V block issued
        ph ^= 1;
      }
    } else {
      // ===== softmax + epilogue: thread = one query row x every second 64-key block =====
      const int quarter = warp & 3;
      const int half = (warp - 2) >> 2;              // key-block residue of this warp: blocks kb = half, half + NP, ...  (0 .. NP-1)
      const int row = quarter * 32 + lane;
      const uint32_t trow = tmem_base + (uint32_t(quarter * 32) << 16);
      uint32_t ph = 0;
      uint32_t eph = 0;   // phase bit per P ring slot (a slot always belongs to the same half: slot parity == block parity)
      const float sc2 = p.scale * 1.4426950408889634f;      // exp(x) = exp2(x * log2 e): one FFMA + MUFU.EX2 per element
      for (int t = 0; t < q_tiles; ++t) {
        long long* tre = (p.trace && blockIdx.x == 0 && warp == 2 && lane == 0 && t < 8) ? p.trace + t * 16 + 4 : nullptr;
        ptx::mbar_wait(s_full, ph);
        ptx::tc_fence_after();
        if (tre) tre[0] = clock64();             // S complete
        // pass 1: row max over this half's key blocks (raw scores; scale > 0 so the max commutes with it).  Both passes walk the 32-column
        // groups of this thread's blocks (kb = half, half + 2, ...; two groups per 64-key block) with double-buffered TMEM loads: group i + 1 is
        // in flight while group i is reduced.  (First version: load, wait, reduce -- the ~300-cycle load latency was exposed 14 times per
        // pass and the exp pass took 7.4 k cycles per 128-query tile, 70 % of the kernel: profiles/r02c_match_trace.txt.)
        uint32_t rr[2][32];
        const int nblk = nkb > half ? (nkb - half + NP - 1) / NP : 0;      // 64-key blocks of this warp
        // Rows of this warp beyond the slot's keypoint count (the last query tile of a 400-keypoint slot has 16 valid rows of 128): the warp
        // keeps the barrier / P-slot protocol going but skips the TMEM reads, the exponentials and the shared-memory writes.  Whatever the
        // P.V MMA then reads for those rows only reaches output rows that are never stored.
        const bool wact = ((part + t * p.q_split) * 128 + quarter * 32) < nq;
        auto col_of = [&](int i) { return (half + NP * (i >> 1)) * 64 + (i & 1) * 32; };
        float mx0 = -INFINITY, mx1 = -INFINITY;
        auto max_group = [&](const uint32_t (&r)[32], int c) {
          if (c + 32 <= nk) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) { mx0 = fmaxf(mx0, __uint_as_float(r[i])); mx1 = fmaxf(mx1, __uint_as_float(r[i + 1])); }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx0 = fmaxf(mx0, (c + i < nk) ? __uint_as_float(r[i]) : -INFINITY);
          }
        };
        {
          int n_it = wact ? 2 * nblk : 0;
          while (n_it > 0 && col_of(n_it - 1) >= nk) --n_it;                 // groups entirely beyond the last key hold no scores
          if constexpr (NP == 2) {
            if (n_it > 0) ptx::tmem_ld32(trow + col_of(0), rr[0]);
            for (int i = 0; i < n_it; i += 2) {
              ptx::tmem_ld_wait();
              if (i + 1 < n_it) ptx::tmem_ld32(trow + col_of(i + 1), rr[1]);
              max_group(rr[0], col_of(i));
              if (i + 1 < n_it) {
                ptx::tmem_ld_wait();
                if (i + 2 < n_it) ptx::tmem_ld32(trow + col_of(i + 2), rr[0]);
                max_group(rr[1], col_of(i + 1));
              }
            }
          } else {
            for (int i = 0; i < n_it; ++i) {
              ptx::tmem_ld32(trow + col_of(i), rr[0]);
              ptx::tmem_ld_wait();
              max_group(rr[0], col_of(i));
            }
          }
        }
        sRed[half * 128 + row] = fmaxf(mx0, mx1);
        if (NP == 2 && wact && nblk > 0) ptx::tmem_ld32(trow + col_of(0), rr[0]);       // first group of pass 2: in flight across the barrier
        if constexpr (NP == 2) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 1, 512;" ::: "memory");
        float mall = fmaxf(sRed[row], sRed[128 + row]);
        if constexpr (NP == 4) mall = fmaxf(mall, fmaxf(sRed[256 + row], sRed[384 + row]));
        const float m2 = mall * sc2;
        if (tre) tre[1] = clock64();             // pass 1 (row max) done
        // pass 2: e = exp(s - max) once per element: accumulate the row sum in fp32 (four partial sums) and hand fp16(e) to the tensor core
        // through shared memory (K-major SWIZZLE_128B, 64 keys per block); the 1/sum normalisation is applied to O in the epilogue.
        float sm[4] = {0.f, 0.f, 0.f, 0.f};
        auto exp_group = [&](const uint32_t (&r)[32], int i) {
          const int kb = half + NP * (i >> 1), hc = i & 1;
          const int ps = kb & (kAttnPSlots - 1);
          if (hc == 0) ptx::mbar_wait(&p_empty[ps], ((eph >> ps) & 1) ^ 1);     // slot free (first use of a fresh barrier passes immediately)
          uint8_t* prow = sP + ps * 16384 + row * 128;
          const int c0 = kb * 64 + hc * 32;
          float e[32];
          if (c0 + 32 <= nk) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { e[j] = ptx::ex2_approx(fmaf(__uint_as_float(r[j]), sc2, -m2)); sm[j & 3] += e[j]; }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { e[j] = (c0 + j < nk) ? ptx::ex2_approx(fmaf(__uint_as_float(r[j]), sc2, -m2)) : 0.f; sm[j & 3] += e[j]; }
          }
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {       // 4 chunks of 8 keys (16 bytes)
            uint32_t pk[4];
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2) {
              __half2 h2 = __floats2half2_rn(e[ch * 8 + q2 * 2], e[ch * 8 + q2 * 2 + 1]);
              pk[q2] = *reinterpret_cast<uint32_t*>(&h2);
            }
            const int chunk = hc * 4 + ch;
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          if (hc == 1) {
            ptx::fence_proxy_async();                // generic-proxy smem writes -> visible to the tensor core (async proxy)
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[ps]);
            eph ^= 1u << ps;
          }
        };
        {
          const int n_it = wact ? 2 * nblk : 0;                              // every 64-key block is written completely (keys >= nk as zeros)
          if (!wact) {
            for (int b2 = 0; b2 < nblk; ++b2) {                              // protocol only: wait for the slot, hand it over unwritten
              const int ps = (half + NP * b2) & (kAttnPSlots - 1);
              ptx::mbar_wait(&p_empty[ps], ((eph >> ps) & 1) ^ 1);
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(&p_full[ps]);
              eph ^= 1u << ps;
            }
          }
          if constexpr (NP == 2) {
            for (int i = 0; i < n_it; i += 2) {
              ptx::tmem_ld_wait();
              ptx::tmem_ld32(trow + col_of(i + 1), rr[1]);
              exp_group(rr[0], i);
              ptx::tmem_ld_wait();
              if (i + 2 < n_it) ptx::tmem_ld32(trow + col_of(i + 2), rr[0]);
              exp_group(rr[1], i + 1);
            }
          } else {
            for (int i = 0; i < n_it; ++i) {
              ptx::tmem_ld32(trow + col_of(i), rr[0]);
              ptx::tmem_ld_wait();
              exp_group(rr[0], i);
            }
          }
        }
        const float sum = (sm[0] + sm[1]) + (sm[2] + sm[3]);
        sRed[NP * 128 + half * 128 + row] = sum;
        if constexpr (NP == 2) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 1, 512;" ::: "memory");
        float sall = sRed[NP * 128 + row] + sRed[NP * 128 + 128 + row];        // the same order in every warp of the row
        if constexpr (NP == 4) sall += sRed[NP * 128 + 256 + row] + sRed[NP * 128 + 384 + row];
        const float inv = 1.f / sall;
        if (tre) tre[2] = clock64();             // pass 2 (exp, P blocks) done
        // epilogue: O (128 x 64 fp32, TMEM columns 0..63) -> fp16 context rows; this warp writes columns half*32 .. +32
        ptx::mbar_wait(o_full, ph);
        ptx::tc_fence_after();
        if (tre) tre[3] = clock64();             // O complete
        const int q = (part + t * p.q_split) * 128 + row;
        __half* o = p.ctx + ((p.row_off ? (long long)q_row0 : (long long)slot * p.cap) + q) * 256 + head * 64;
        if constexpr (NP == 2) {
          const int c = half * 32;
          uint32_t r[32];
          ptx::tmem_ld32(trow + c, r);
          ptx::tmem_ld_wait();
          if (q < nq) {
#pragma unroll
            for (int i = 0; i < 32; i += 16) {
              uint32_t h[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                __half2 h2 = __floats2half2_rn(__uint_as_float(r[i + 2 * e]) * inv, __uint_as_float(r[i + 2 * e + 1]) * inv);
                h[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              ptx::st_global_256(o + c + i, h);       // ctx rows are 512 B, head * 64 + c + i a multiple of 16 halves: 32-byte aligned
            }
          }
        } else {
          const int c = half * 16;
          uint32_t r[16];
          ptx::tmem_ld16(trow + c, r);
          ptx::tmem_ld_wait();
          if (q < nq) {
            uint32_t h[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              __half2 h2 = __floats2half2_rn(__uint_as_float(r[2 * e]) * inv, __uint_as_float(r[2 * e + 1]) * inv);
              h[e] = *reinterpret_cast<uint32_t*>(&h2);
            }
            ptx::st_global_256(o + c, h);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(s_free);
        if (tre) tre[4] = clock64();             // context stored
        ph ^= 1;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

}  // namespace airfe
