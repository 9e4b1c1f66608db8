// tcgen05 3x3 convolution with HALO REUSE for sm_100a  (the dominant kernel of the path).
//
// The generic kernel (tc_gemm.cuh) streams one shifted 128-pixel A tile per tap: 9x the activation bytes through L2->SMEM,
// which bounds small-C layers at ~350 TFLOP/s.  Here a CTA loads ONE halo tile of (8S+2) x 18 pixels x 64 channels per K block
// and the nine taps are nine tcgen05 shared-memory descriptors into that same tile:
//     start address = halo + ((ky*HW + kx + 8*s) * 128 B),  SBO = HW * 128 B   (HW = 8S + 2 halo pixels per row)
// i.e. MMA row group g (8 consecutive pixels of image row y0+g) sits 1 halo row further down.  This relies on the measured
// fact (profiles/r01_umma_descriptor_probe.txt) that SWIZZLE_128B is applied on absolute shared-memory address bits, so
// a descriptor may start at any 128-byte row and use any SBO.
//   * S = 1 or 2 strips of 8x16 pixels share the halo tile and each weight tile (S accumulators in TMEM, double-buffered).
//   * weights: streamed per (K block, tap) through their own ring, or -- when 9*C*N*2 bytes fit -- resident in shared memory
//     for the whole persistent CTA (64->64, 64->32, 32->32: the 512x512 layers).
//   * epilogue: +bias, ReLU, fp16 NHWC store and/or fused 2x2 max-pool (warp shuffles: pool partners are lane^1 and lane^8).
#pragma once
#include "ptx.cuh"

namespace airfe {

struct ConvParams {
  CUtensorMap tmA;  // 4-D (C, W, H, B), box (64, 8S+2, 18, 1), SWIZZLE_128B
  CUtensorMap tmB;  // 4-D (K, N, 1, 1), box (64, block_n, 1, 1)
  int kblocks, c_in_pad;
  int kw;           // channels per K block: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows, for C_in = 32 layers)
  int strips;
  int tiles_x, tiles_y, B, W, H;
  int n_tiles, block_n, n_valid;
  int b_resident;
  int fold;         // 1 / 2: kx taps folded into N (tc_conv3x3_fold.cuh; 2 = sixteen epilogue warps); tiles are 14 output columns x 16 rows
  int stages_a, stages_b;
  int prewait;      // MMA issuer polls the barriers of unit u + 1 before it issues unit u (0: after; AIRFE_PREWAIT=1 switches it on, for A/B timing)
  int nacc;         // TMEM accumulator buffers (2..4): deeper than 2 hides the MMA -> epilogue -> MMA hand-shake latency on small-N layers
  const float* bias;
  int relu;
  __half* out;        // full-resolution fp16 NHWC store (nullptr: skip)
  long long out_sb, out_sy, out_sx;
  int up2;            // 1: `out` is the nearest-neighbour 2x up-sampled map (G2 hourglass decoder, Resize nodes before the `deconv` convs): pixel (y, x)
                      // is stored at (2y, 2x), (2y, 2x+1), (2y+1, 2x), (2y+1, 2x+1); strides are those of the up-sampled buffer (32-byte aligned rows of 16 channels)
  __half* pool_out;   // fused 2x2/2 max-pool store (nullptr: skip)
  long long pool_sb, pool_sy, pool_sx;
  long long* trace;   // authoring aid (airfe_debug_conv_trace): CTA 0 writes clock64 stamps of its first 64 tiles, 8 slots per tile
};

constexpr int kConvThreads = 320;   // warp0 TMA, warp1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kConvWideThreads = 576;   // EW = 4: warps 2-17 epilogue (four per TMEM lane quarter = 4.5 warps per scheduler instead of 2.5)
constexpr int kConvTH = 16;

__host__ __device__ constexpr int conv_a_bytes(int strips, int kw = 64) { return ((8 * strips + 2) * (kConvTH + 2) * kw * 2 + 1023) / 1024 * 1024; }
__host__ __device__ inline int conv_b_bytes(int block_n, int kw = 64) { return (block_n * kw * 2 + 1023) / 1024 * 1024; }
__host__ __device__ inline int conv_acc_stride(int block_n) { return (block_n + 31) / 32 * 32; }
__host__ __device__ inline int conv_nacc(int block_n, int strips) {
  int n = 512 / (strips * conv_acc_stride(block_n));
  return n > 4 ? 4 : (n < 2 ? 2 : n);
}
__host__ __device__ inline int conv_tmem_cols(int block_n, int strips) {
  int need = conv_nacc(block_n, strips) * strips * conv_acc_stride(block_n), c = 32;
  while (c < need) c <<= 1;
  return c;
}

// Template parameters fix everything the MMA issue loop would otherwise branch on: measured with the clock64 trace
// (tools/trace_conv.py, profiles/r01_conv_trace.txt) the runtime-generic loop spent ~550 cycles of uniform-datapath
// instructions per tap, i.e. the single issuing warp -- not the tensor pipe (48 cycles per 128x64x16 MMA in SS mode,
// tools/probe_mma_rate.cu) -- bounded every small-N layer.
// EW = epilogue warps per TMEM lane quarter (2 or 4).  The epilogue of the small-N / memory-heavy layers is latency-bound with two warps per
// scheduler (ncu: 65 % of the cycles no warp eligible, stalls = fixed-latency waits + shared-memory scoreboard); EW = 4 doubles the warps
// that hide each other's latencies at the same per-element arithmetic (results are bit-identical: a chunk of 16 channels of one pixel is
// always processed by one thread, only the assignment of chunks / strips to warps changes).
template <int KW, int STRIPS, bool BRES, int EW = 2>
__global__ void __launch_bounds__(64 + 128 * EW, 1) tc_conv3x3_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int a_bytes = conv_a_bytes(STRIPS, KW);
  const int b_bytes = conv_b_bytes(p.block_n, KW);
  constexpr int rowb = KW * 2;                    // bytes per pixel row of a K block
  const int n_b_slots = BRES ? 9 * p.kblocks : p.stages_b;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.stages_a * a_bytes;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(smem_b + n_b_slots * b_bytes);
  uint64_t* empty_a = full_a + p.stages_a;
  uint64_t* full_b = empty_a + p.stages_a;          // [stages_b] (streaming) or [1] (resident: all weights landed)
  uint64_t* empty_b = full_b + (BRES ? 1 : p.stages_b);
  uint64_t* tmem_full = empty_b + (BRES ? 1 : p.stages_b);
  uint64_t* tmem_empty = tmem_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int HW = 8 * STRIPS + 2;
  const int m_tiles = p.tiles_x * p.tiles_y * p.B;
  const int total_tiles = m_tiles * p.n_tiles;
  const uint32_t acc_stride = conv_acc_stride(p.block_n);
  const uint32_t tmem_cols = conv_tmem_cols(p.block_n, STRIPS);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmB);
    for (int s = 0; s < p.stages_a; ++s) { ptx::mbar_init(&full_a[s], 1); ptx::mbar_init(&empty_a[s], 1); }
    const int nb = BRES ? 1 : p.stages_b;
    for (int s = 0; s < nb; ++s) { ptx::mbar_init(&full_b[s], 1); ptx::mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < 4; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 4 * EW); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, tmem_cols); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  __shared__ __align__(16) float s_bias[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s_bias[i] = (p.bias && i < p.n_valid) ? p.bias[i] : 0.f;   // weights: not produced by a kernel
  __syncthreads();
  ptx::pdl_launch_dependents();   // the next kernel may start its own prologue on SMs this grid has left
  ptx::pdl_wait();                // everything above overlapped the predecessor's tail; activations are touched only below

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      if (BRES) {   // weights of this CTA's (single) N tile: load once, keep for every tile
        ptx::mbar_arrive_expect_tx(&full_b[0], (uint32_t)(9 * p.kblocks * p.block_n * rowb));
        for (int cb = 0; cb < p.kblocks; ++cb)
          for (int tap = 0; tap < 9; ++tap)
            ptx::tma_load_4d(smem_b + (cb * 9 + tap) * b_bytes, &p.tmB, &full_b[0], tap * p.c_in_pad + cb * KW, 0, 0, 0);
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % p.n_tiles, mt = t / p.n_tiles;
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tz = mt / (p.tiles_x * p.tiles_y);
        const int x0 = tx * 8 * STRIPS, y0 = ty * kConvTH;
        for (int cb = 0; cb < p.kblocks; ++cb) {
          ptx::mbar_wait(&empty_a[sa], pa ^ 1);
          if (p.trace && blockIdx.x == 0 && cb == 0 && t / (int)gridDim.x < 64) p.trace[(t / gridDim.x) * 8 + 0] = clock64();
          ptx::mbar_arrive_expect_tx(&full_a[sa], (uint32_t)(HW * (kConvTH + 2) * rowb));
          ptx::tma_load_4d(smem_a + sa * a_bytes, &p.tmA, &full_a[sa], cb * KW, x0 - 1, y0 - 1, tz);
          if (++sa == p.stages_a) { sa = 0; pa ^= 1; }
          if (!BRES) {
            for (int tap = 0; tap < 9; ++tap) {
              ptx::mbar_wait(&empty_b[sb], pb ^ 1);
              ptx::mbar_arrive_expect_tx(&full_b[sb], (uint32_t)(p.block_n * rowb));
              ptx::tma_load_4d(smem_b + sb * b_bytes, &p.tmB, &full_b[sb], tap * p.c_in_pad + cb * KW, nt * p.block_n, 0, 0);
              if (++sb == p.stages_b) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the (uniform) loop, one elected lane issues =====
    const uint32_t idesc = ptx::make_idesc_f16(128, p.block_n, 0);
    int sa = 0, sb = 0, acc = 0;
    uint32_t pa = 0, pb = 0, acc_phase = 0;
    if (BRES) { ptx::mbar_wait(&full_b[0], 0); ptx::tc_fence_after(); }
    constexpr int row16 = rowb >> 4;                // pixel-row pitch in 16-byte units (8 or 4)
    constexpr int ksteps = KW >> 4;                 // tcgen05.mma K = 16 steps per block (4 or 2)
    const uint64_t da_const = KW == 64 ? ptx::smem_desc_base_sw128((uint32_t)HW * 128) : ptx::smem_desc_base_sw64((uint32_t)HW * 64);
    const uint64_t db_const = KW == 64 ? ptx::smem_desc_base_sw128(1024) : ptx::smem_desc_base_sw64(512);
    const uint64_t da_base = da_const + (ptx::smem_u32(smem_a) >> 4);
    const uint64_t db_base = db_const + (ptx::smem_u32(smem_b) >> 4);
    const uint32_t b16 = (uint32_t)b_bytes >> 4;
    if constexpr (BRES) {
      // Resident-weights layers (the 512 x 512 maps, N <= 64: a tile is only 36 / 72 MMAs long).  A clock64 trace of these layers shows
      // ~840 cycles between the last accepted MMA of a tile and the first of the next.  Hypothesis tested in round 2: the issuing thread
      // polls the next tile's barriers only after its (back-pressured) issue block, so polling them one unit EARLIER (p.prewait) would close
      // the gap.  Measured (profiles/r02_prewait_ab.txt): no change -- the period is set by the arrival of the halo tiles (these layers move
      // 2.7 .. 3.8 TB/s of HBM traffic, see profiles/r02*_ncu_all_kernels_cfg2.txt), the wait merely moves.  The option stays for A/B runs.
      const bool prewait = p.prewait && p.stages_a >= 3 && p.nacc >= 2;
      if ((int)blockIdx.x < total_tiles) { ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1); ptx::mbar_wait(&full_a[sa], pa); }
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const bool tr = p.trace && blockIdx.x == 0 && lane == 0 && t / (int)gridDim.x < 64;
        long long* trp = tr ? p.trace + (t / gridDim.x) * 8 : nullptr;
        if (tr) { trp[1] = clock64(); trp[2] = trp[1]; }
        const uint32_t d0 = tmem_base + (uint32_t)(acc * STRIPS) * acc_stride;
        const bool next_tile = t + (int)gridDim.x < total_tiles;
        int nacc = acc + 1;
        uint32_t nphase = acc_phase;
        if (nacc == p.nacc) { nacc = 0; nphase ^= 1; }
        for (int cb = 0; cb < p.kblocks; ++cb) {
          const bool last_cb = cb == p.kblocks - 1;
          int nsa = sa + 1;
          uint32_t npa = pa;
          if (nsa == p.stages_a) { nsa = 0; npa ^= 1; }
          if (prewait) {
            if (!last_cb) ptx::mbar_wait(&full_a[nsa], npa);
            else if (next_tile) { ptx::mbar_wait(&tmem_empty[nacc], nphase ^ 1); ptx::mbar_wait(&full_a[nsa], npa); }
          }
          ptx::tc_fence_after();
          const uint64_t da_stage = da_base + (uint64_t)((uint32_t)(sa * a_bytes) >> 4);
          const uint32_t first = cb != 0;   // accumulate flag of the very first MMA of a tile
          const uint64_t db_cb = db_base + (uint64_t)((uint32_t)(cb * 9) * b16);
          if (ptx::elect_one()) {           // 9 taps x ksteps x STRIPS MMAs as one straight-line issue block under a single election
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint64_t da_tap = da_stage + (uint64_t)(((tap / 3) * HW + tap % 3) * row16);
              const uint64_t db = db_cb + (uint64_t)((uint32_t)tap * b16);
#pragma unroll
              for (int k = 0; k < ksteps; ++k) ptx::umma_f16(d0, da_tap + 2 * k, db + 2 * k, idesc, (tap | k) ? 1u : first);
              if (STRIPS == 2) {
#pragma unroll
                for (int k = 0; k < ksteps; ++k) ptx::umma_f16(d0 + acc_stride, da_tap + 8 * row16 + 2 * k, db + 2 * k, idesc, (tap | k) ? 1u : first);
              }
            }
            ptx::umma_commit(&empty_a[sa]);
            if (last_cb) ptx::umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (!prewait) {
            if (!last_cb) ptx::mbar_wait(&full_a[nsa], npa);
            else if (next_tile) { ptx::mbar_wait(&tmem_empty[nacc], nphase ^ 1); ptx::mbar_wait(&full_a[nsa], npa); }
          }
          sa = nsa; pa = npa;
        }
        if (tr) trp[3] = clock64();
        acc = nacc; acc_phase = nphase;
      }
    } else {
    bool b_primed = false;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const bool tr = p.trace && blockIdx.x == 0 && lane == 0 && t / (int)gridDim.x < 64;
      long long* trp = tr ? p.trace + (t / gridDim.x) * 8 : nullptr;
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      if (tr) trp[1] = clock64();
      const uint32_t d0 = tmem_base + (uint32_t)(acc * STRIPS) * acc_stride;
      for (int cb = 0; cb < p.kblocks; ++cb) {
        ptx::mbar_wait(&full_a[sa], pa);
        ptx::tc_fence_after();
        if (tr && cb == 0) trp[2] = clock64();
        const uint64_t da_stage = da_base + (uint64_t)((uint32_t)(sa * a_bytes) >> 4);
        const uint32_t first = cb != 0;   // accumulate flag of the very first MMA of a tile
        {
          // streamed weights: one (K block, tap) tile per ring stage.  As in the resident-weights path the barrier of the NEXT weight tile is
          // polled before the MMAs of the current one are issued (the ring has >= 3 stages), so consecutive issue blocks are back to back;
          // the first tile of the kernel is waited for here.
          const bool pw = p.prewait && p.stages_b >= 3;
          if (!b_primed) { ptx::mbar_wait(&full_b[sb], pb); b_primed = true; }
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            int nsb = sb + 1;
            uint32_t npb = pb;
            if (nsb == p.stages_b) { nsb = 0; npb ^= 1; }
            const bool more_b = tap < 8 || cb + 1 < p.kblocks || t + (int)gridDim.x < total_tiles;     // another weight tile follows in this CTA
            if (pw && more_b) ptx::mbar_wait(&full_b[nsb], npb);
            ptx::tc_fence_after();
            const uint64_t db = db_base + (uint64_t)((uint32_t)sb * b16);
            const uint64_t da_tap = da_stage + (uint64_t)(((tap / 3) * HW + tap % 3) * row16);
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < ksteps; ++k) ptx::umma_f16(d0, da_tap + 2 * k, db + 2 * k, idesc, (tap | k) ? 1u : first);
              if (STRIPS == 2) {
#pragma unroll
                for (int k = 0; k < ksteps; ++k) ptx::umma_f16(d0 + acc_stride, da_tap + 8 * row16 + 2 * k, db + 2 * k, idesc, (tap | k) ? 1u : first);
              }
              ptx::umma_commit(&empty_b[sb]);
              if (tap == 8) ptx::umma_commit(&empty_a[sa]);
            }
            __syncwarp();
            if (!pw && more_b) ptx::mbar_wait(&full_b[nsb], npb);
            sb = nsb; pb = npb;
          }
        }
        if (++sa == p.stages_a) { sa = 0; pa ^= 1; }
      }
      if (ptx::elect_one()) ptx::umma_commit(&tmem_full[acc]);
      __syncwarp();
      if (tr) trp[3] = clock64();
      if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
    }
    }
  } else {
    // ===== epilogue: 4 * EW warps; warp (2 + e) owns TMEM lane quarter (warp & 3) and one part of the work (strip x column part) =====
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;                             // 0 .. EW - 1
    const int lx = lane & 7, ly = quarter * 4 + (lane >> 3);      // pixel of this lane inside an 8 x 16 strip
    const int chunks = p.block_n / 16;
    constexpr int CP = (STRIPS == 2) ? EW / 2 : EW;               // column parts per strip
    const int s_mine = (STRIPS == 2) ? (part & 1) : 0;
    const int cpart = (STRIPS == 2) ? (part >> 1) : part;
    const int c_begin = (chunks * cpart + CP - 1) / CP;
    const int c_end = (chunks * (cpart + 1) + CP - 1) / CP;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int nt = t % p.n_tiles, mt = t / p.n_tiles;
      const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tz = mt / (p.tiles_x * p.tiles_y);
      const int y = ty * kConvTH + ly;
      const int x = (tx * STRIPS + s_mine) * 8 + lx;
      const bool valid = (x < p.W) && (y < p.H);
      const int n0 = nt * p.block_n;
      __half* o_full = p.out ? p.out + (long long)tz * p.out_sb + (long long)(y << p.up2) * p.out_sy + (long long)(x << p.up2) * p.out_sx : nullptr;
      __half* o_pool = p.pool_out ? p.pool_out + (long long)tz * p.pool_sb + (long long)(y >> 1) * p.pool_sy + (long long)(x >> 1) * p.pool_sx : nullptr;
      const bool pool_lane = valid && !(lane & 1) && !(lane & 8);
      const bool tr = p.trace && blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 1 + 4 * EW) && t / (int)gridDim.x < 64;
      long long* trp = tr ? p.trace + (t / gridDim.x) * 8 + (warp == 2 ? 4 : 6) : nullptr;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      if (tr) trp[0] = clock64();
      const uint32_t taddr = tmem_base + (uint32_t)(acc * STRIPS + s_mine) * acc_stride + (uint32_t(quarter * 32) << 16);
      uint32_t r[2][16];
      if (c_begin < c_end) ptx::tmem_ld16(taddr + c_begin * 16, r[0]);
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int cc = c + u;
          if (cc >= c_end) break;
          ptx::tmem_ld_wait();                                              // chunk cc has landed in r[u]
          if (cc + 1 < c_end) ptx::tmem_ld16(taddr + (cc + 1) * 16, r[u ^ 1]);   // next chunk in flight while this one is processed
          const int nbase = n0 + cc * 16;
          uint32_t h[8];
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + nbase);     // 4 x LDS.128 instead of 16 scalar loads
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bb = b4[i];
            float f0 = __uint_as_float(r[u][4 * i]) + bb.x, f1 = __uint_as_float(r[u][4 * i + 1]) + bb.y;
            float f2 = __uint_as_float(r[u][4 * i + 2]) + bb.z, f3 = __uint_as_float(r[u][4 * i + 3]) + bb.w;
            if (p.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f); }
            __half2 ha = __floats2half2_rn(f0, f1), hb = __floats2half2_rn(f2, f3);
            h[2 * i] = *reinterpret_cast<uint32_t*>(&ha);
            h[2 * i + 1] = *reinterpret_cast<uint32_t*>(&hb);
          }
          if (o_full && valid && nbase < p.n_valid) {
            __half* o = o_full + nbase;
            if (nbase + 16 <= p.n_valid) {
              if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {
                ptx::st_global_256(o, h);
                if (p.up2) {     // fused nearest 2x up-sampling (the host admits only whole, 32-byte aligned chunks): the other three copies of this pixel
                  ptx::st_global_256(o + p.out_sx, h);
                  ptx::st_global_256(o + p.out_sy, h);
                  ptx::st_global_256(o + p.out_sy + p.out_sx, h);
                }
              } else {
                *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(o + 8) = make_uint4(h[4], h[5], h[6], h[7]);
              }
            } else {
              unsigned short* os = reinterpret_cast<unsigned short*>(o);   // static indices only: h[] must stay in registers
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (nbase + 2 * i < p.n_valid) os[2 * i] = (unsigned short)(h[i] & 0xffffu);
                if (nbase + 2 * i + 1 < p.n_valid) os[2 * i + 1] = (unsigned short)(h[i] >> 16);
              }
            }
          }
          if (p.pool_out) {
            // 2x2 max-pool on the fp16-rounded values (identical to pooling the stored tensor): partners are lane^1 (x) and lane^8 (y)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __half2 a = *reinterpret_cast<__half2*>(&h[i]);
              uint32_t o1 = __shfl_xor_sync(0xffffffffu, h[i], 1);
              a = __hmax2(a, *reinterpret_cast<__half2*>(&o1));
              uint32_t cur = *reinterpret_cast<uint32_t*>(&a);
              uint32_t o8 = __shfl_xor_sync(0xffffffffu, cur, 8);
              a = __hmax2(a, *reinterpret_cast<__half2*>(&o8));
              h[i] = *reinterpret_cast<uint32_t*>(&a);
            }
            if (pool_lane && nbase < p.n_valid) {
              __half* o = o_pool + nbase;
              if (nbase + 16 <= p.n_valid) {
                if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {
                  ptx::st_global_256(o, h);
                } else {
                  *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
                  *reinterpret_cast<uint4*>(o + 8) = make_uint4(h[4], h[5], h[6], h[7]);
                }
              } else {
                unsigned short* os = reinterpret_cast<unsigned short*>(o);   // static indices only: h[] must stay in registers
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if (nbase + 2 * i < p.n_valid) os[2 * i] = (unsigned short)(h[i] & 0xffffu);
                  if (nbase + 2 * i + 1 < p.n_valid) os[2 * i + 1] = (unsigned short)(h[i] >> 16);
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (tr) trp[1] = clock64();
      if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, tmem_cols); }
}

}  // namespace airfe
