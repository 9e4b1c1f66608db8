// tcgen05 3x3 convolution with the three kx taps FOLDED INTO THE N DIMENSION (small-N layers: C_out = 32 / 64, C_in <= 64).
//
// Why.  With both operands in shared memory a 128 x N x 16 tcgen05.mma costs max(N/2, 32 + N/4) cycles (tools/probe_mma_rate.cu,
// profiles/r01_mma_issue_rate_probe.txt): the 4 KB A slice is re-read at 128 B/clk for every instruction, so the nine-tap halo kernel
// (tc_conv3x3.cuh: nine MMAs of N = 32 / 64 per K step) cannot exceed 40 % / 67 % of the tensor peak on the 512 x 512 layers.
// Here ONE MMA per (ky, K step) multiplies the A slice with the weights of all three kx taps side by side (N' = 3N = 96 / 192 columns):
//     P_kx[y][x] = sum_{ky, c} in[y - 1 + ky][x][c] * w[ky][kx][c][n]          (x = input column, NOT shifted)
//     out[y][x]  = P_0[y][x - 1] + P_1[y][x] + P_2[y][x + 1]                     (+ bias, ReLU)
// so each A slice is read once per three taps: N' = 96 -> 56 cycles for what took 120, N' = 192 -> 96 for what took 144.  The +-1 column
// shift happens in the epilogue with two warp shuffles per value: an MMA strip is 16 input columns x 8 rows (M = 128), TMEM lane
// l of warp quarter q holds column (l & 15) of strip row 2q + (l >> 4), so the column neighbours are lane -+ 1 and a warp produces
// 14 valid output columns per row (tiles advance by 14 columns; 12.5 % of the MMA rows are halo) and the 2x2 max-pool partners are
// lane + 1 and lane ^ 16.  The halo tile (16 columns x (8S + 2) rows x KW channels, one 4-D TMA box, out-of-bounds zero fill = the
// convolution padding) is a dense K-major SWIZZLE_128B / 64B matrix: strip s / tap row ky is the canonical 128-row operand that starts
// (8s + ky) * 16 pixel rows into it (a multiple of 1 KiB, SBO = 8 rows: no reliance on unaligned descriptor starts).
// Weights stay resident in shared memory in tap order, so the three kx blocks of one ky are one contiguous 3N-row B operand.
// Accumulators: one TMEM buffer of 3N columns per STRIP in a ring (2 buffers for N = 64, 4 for N = 32): the epilogue of strip i overlaps
// the MMAs of strip i + 1.
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2-9: epilogue (quarter = warp & 3, channel half = (warp - 2) >> 2)
#pragma once
#include "tc_conv3x3.cuh"

namespace airfe {

constexpr int kFoldTX = 14;      // valid output columns per tile (16 loaded)
constexpr int kFoldStrips = 2;   // strips of 8 rows per halo tile

__host__ __device__ constexpr int fold_a_bytes(int kw, int strips = kFoldStrips) { return 16 * (8 * strips + 2) * kw * 2; }   // one K block of a halo tile: 36 KiB (KW = 64, 2 strips), 18 KiB (KW = 32), 20 KiB (KW = 64, 1 strip)
__host__ __device__ constexpr int fold_nbuf(int n) { return n == 64 ? 2 : 4; }

constexpr int kFoldWideThreads = 576;   // WIDE: warps 2-17 = sixteen epilogue warps (four per TMEM lane quarter / scheduler)

// WIDE = false: 8 epilogue warps, each software-pipelines its (strip, chunk) items (register double buffer).
// WIDE = true : 16 epilogue warps, one (strip, chunk group) per warp and tile, no intra-warp pipelining: the epilogue of this kernel is
//               issue / latency bound (ncu: 2.5 resident warps per scheduler, issue slot 47 % busy, stalls = fixed-latency waits and the
//               shuffle / store scoreboard), which more resident warps hide directly.
// S  = strips of 8 rows per halo tile (2; 1 for the two-K-block variant, whose halo tile is twice as deep)
// KB = K blocks of 64 input channels (1; 2 = C_in 128: both K blocks of a tile form one stage and a strip accumulates over 2 x 3 x 4 MMAs).
//      With KB = 2 the layer's 64 output channels are TWO resident N tiles of 32 (p.n_tiles = 2; a CTA keeps the N tile blockIdx.x % 2, the grid
//      is even): 9 x 128 x 64 fp16 weights (147 KB) do not fit next to the halo ring, 9 x 128 x 32 do.  WIDE only.
template <int KW, int N, bool WIDE, int S = kFoldStrips, int KB = 1>
__global__ void __launch_bounds__(WIDE ? kFoldWideThreads : kConvThreads, 1) tc_conv3x3_fold_kernel(const __grid_constant__ ConvParams p) {
  static_assert(WIDE || (S == kFoldStrips && KB == 1), "the 8-warp epilogue exists for the one-K-block, two-strip shape only");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int a_one = fold_a_bytes(KW, S);   // one K block of the halo tile
  constexpr int a_bytes = KB * a_one;          // ring stage
  constexpr int rowb = KW * 2;                 // bytes per pixel row
  constexpr int b_bytes = N * rowb;            // one tap's weights: N rows (a multiple of 1 KiB for every instantiation)
  constexpr int NB = fold_nbuf(N);
  constexpr int ACC = 3 * N;                   // TMEM columns per strip
  static_assert(b_bytes % 1024 == 0 && a_one % 1024 == 0, "operand blocks must keep the 1 KiB swizzle phase");
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.stages_a * a_bytes;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(smem_b + 9 * KB * b_bytes);
  uint64_t* empty_a = full_a + p.stages_a;
  uint64_t* full_b = empty_a + p.stages_a;
  uint64_t* tmem_full = full_b + 1;
  uint64_t* tmem_empty = tmem_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_x * p.tiles_y * p.B * p.n_tiles;      // t = m tile * n_tiles + n tile; the N tile of a CTA never changes
  const int nt0 = (int)blockIdx.x % p.n_tiles;
  constexpr uint32_t tmem_cols = 512;          // 2 x 192 or 4 x 96 columns: the next power of two

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmB);
    for (int s = 0; s < p.stages_a; ++s) { ptx::mbar_init(&full_a[s], 1); ptx::mbar_init(&empty_a[s], 1); }
    ptx::mbar_init(&full_b[0], 1);
    for (int s = 0; s < 4; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 8); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, tmem_cols); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  __shared__ __align__(16) float s_bias[64];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_bias[i] = (p.bias && i < N && nt0 * N + i < p.n_valid) ? p.bias[nt0 * N + i] : 0.f;   // weights: not produced by a kernel
  __syncthreads();
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();                // activations are touched only below

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: the nine weight taps once, then one halo tile per output tile =====
      ptx::mbar_arrive_expect_tx(&full_b[0], (uint32_t)(9 * KB * b_bytes));
      for (int cb = 0; cb < KB; ++cb)
        for (int tap = 0; tap < 9; ++tap) ptx::tma_load_4d(smem_b + (cb * 9 + tap) * b_bytes, &p.tmB, &full_b[0], tap * p.c_in_pad + cb * KW, nt0 * N, 0, 0);
      int sa = 0;
      uint32_t pa = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int mt = t / p.n_tiles;
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tz = mt / (p.tiles_x * p.tiles_y);
        ptx::mbar_wait(&empty_a[sa], pa ^ 1);
        if (p.trace && blockIdx.x == 0 && t / (int)gridDim.x < 64) p.trace[(t / gridDim.x) * 8 + 0] = clock64();
        ptx::mbar_arrive_expect_tx(&full_a[sa], (uint32_t)a_bytes);
#pragma unroll
        for (int cb = 0; cb < KB; ++cb) ptx::tma_load_4d(smem_a + sa * a_bytes + cb * a_one, &p.tmA, &full_a[sa], cb * KW, tx * kFoldTX - 1, ty * 8 * S - 1, tz);
        if (++sa == p.stages_a) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: per strip 3 (ky) x KW/16 MMAs of 128 x 3N x 16 as one straight-line block under a single election =====
    const uint32_t idesc = ptx::make_idesc_f16(128, 3 * N, 0);
    constexpr int ksteps = KW >> 4;
    const uint64_t d_const = KW == 64 ? ptx::smem_desc_base_sw128(1024) : ptx::smem_desc_base_sw64(512);
    const uint64_t da_base = d_const + (ptx::smem_u32(smem_a) >> 4);
    const uint64_t db_base = d_const + (ptx::smem_u32(smem_b) >> 4);
    constexpr uint32_t row16 = (16 * rowb) >> 4;   // one halo row (16 pixels) in 16-byte units
    constexpr uint32_t b16 = (uint32_t)b_bytes >> 4;
    int sa = 0, buf = 0;
    uint32_t pa = 0, bphase = 0;
    ptx::mbar_wait(&full_b[0], 0);
    ptx::tc_fence_after();
    if ((int)blockIdx.x < total_tiles) { ptx::mbar_wait(&full_a[sa], pa); ptx::mbar_wait(&tmem_empty[buf], bphase ^ 1); }
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const bool next_tile = t + (int)gridDim.x < total_tiles;
      long long* trp = (p.trace && blockIdx.x == 0 && lane == 0 && t / (int)gridDim.x < 64) ? p.trace + (t / gridDim.x) * 8 : nullptr;
      if (trp) trp[1] = clock64();
      const uint64_t da_stage = da_base + (uint64_t)((uint32_t)(sa * a_bytes) >> 4);
      int nsa = sa + 1;
      uint32_t npa = pa;
      if (nsa == p.stages_a) { nsa = 0; npa ^= 1; }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * ACC);
        if (ptx::elect_one()) {
#pragma unroll
          for (int cb = 0; cb < KB; ++cb)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint64_t da = da_stage + (uint64_t)((uint32_t)(cb * (a_one >> 4)) + (uint32_t)(8 * s + ky) * row16);
              const uint64_t db = db_base + (uint64_t)((uint32_t)(cb * 9 + 3 * ky) * b16);
#pragma unroll
              for (int k = 0; k < ksteps; ++k) ptx::umma_f16(d0, da + 2 * k, db + 2 * k, idesc, (cb | ky | k) ? 1u : 0u);
            }
          if (s == S - 1) ptx::umma_commit(&empty_a[sa]);
          ptx::umma_commit(&tmem_full[buf]);
        }
        __syncwarp();
        if (trp && s == 0) trp[2] = clock64();
        if (++buf == NB) { buf = 0; bphase ^= 1; }
        // poll the next unit's barriers AFTER issuing (the issue block above is back-pressured by the tensor pipe anyway)
        if (s < S - 1) {
          ptx::mbar_wait(&tmem_empty[buf], bphase ^ 1);
        } else if (next_tile) {
          ptx::mbar_wait(&full_a[nsa], npa);
          ptx::mbar_wait(&tmem_empty[buf], bphase ^ 1);
        }
      }
      if (trp) trp[3] = clock64();
      sa = nsa; pa = npa;
    }
  } else if constexpr (WIDE) {
    // ===== epilogue, 16 warps: warp -> TMEM lane quarter (warp & 3), strip of the tile ((warp - 2) >> 3) and chunk group (((warp - 2) >> 2) & 1) =====
    // This epilogue is instruction-issue bound (one item = 16 channels x 32 pixels costs ~160 essential instructions: 3 TMEM loads, 32
    // shuffles, 48 adds, ReLU, packing, the store), so everything else is kept out of the per-tile path: the tile coordinates advance
    // incrementally (no division), the thread's store offset inside a tile is computed once, and the host only selects this kernel for
    // layers whose channel count equals N and whose rows are 32-byte aligned (no partial-chunk store path).
    const int quarter = warp & 3;
    const int g = (warp - 2) >> 2;                     // 0..3: chunk group g & 1, strip-sequence parity g >> 1
    // The CTA's strips form one sequence gi = tile iteration * S + strip; this warp serves the strips with gi % 2 == g >> 1:
    // S = 2: strip (g >> 1) of every tile; S = 1: every second tile.
    constexpr int iter_step = (S == 1) ? 2 : 1;
    const int s_mine = (S == 2) ? (g >> 1) : 0;
    const int first_iter = (S == 1) ? (g >> 1) : 0;
    constexpr int CPW = N / 32;                        // chunks of 16 output channels per warp: 1 (N = 32) or 2 (N = 64)
    const int c_begin = (g & 1) * CPW;
    const int xx = lane & 15;
    const int yr = 8 * s_mine + 2 * quarter + (lane >> 4);          // row inside the 8S-row tile
    const uint32_t lane_sel = uint32_t(quarter * 32) << 16;
    const bool x_ok = (xx >= 1) && (xx <= kFoldTX);
    const bool pool_sel = (xx & 1) && !(lane & 16);                 // x even (tiles start at even columns), y even
    const long long off_full = (long long)yr * p.out_sy + (long long)(xx - 1) * p.out_sx + nt0 * N + c_begin * 16;
    const long long off_pool = (long long)(yr >> 1) * p.pool_sy + (long long)((xx - 1) >> 1) * p.pool_sx + nt0 * N + c_begin * 16;
    // m-tile coordinates of this warp's first tile, then += (m tiles per step) with carries: no division on the per-tile path
    const int txy = p.tiles_x * p.tiles_y;
    const int gm = (int)gridDim.x / p.n_tiles;         // m-tile stride of one CTA iteration (the grid is a multiple of n_tiles)
    const int m0 = (int)blockIdx.x / p.n_tiles + first_iter * gm, dm = iter_step * gm;
    int tz = m0 / txy, ty = (m0 % txy) / p.tiles_x, tx = m0 % p.tiles_x;
    const int dz = dm / txy, dy = (dm % txy) / p.tiles_x, dx = dm % p.tiles_x;
    int gi = first_iter * S + s_mine;                  // position of this warp's strip in the CTA's strip sequence: buffer gi % NB, phase (gi / NB) & 1
    for (int t = blockIdx.x + first_iter * gridDim.x; t < total_tiles; t += iter_step * gridDim.x, gi += 2) {
      const int x = tx * kFoldTX - 1 + xx;
      const int y = ty * (8 * S) + yr;
      const bool valid = x_ok && (x < p.W) && (y < p.H);
      __half* o_full = p.out + ((long long)tz * p.out_sb + (long long)(ty * (8 * S)) * p.out_sy + (long long)(tx * kFoldTX) * p.out_sx + off_full);
      __half* o_pool = p.pool_out + ((long long)tz * p.pool_sb + (long long)(ty * (4 * S)) * p.pool_sy + (long long)(tx * (kFoldTX / 2)) * p.pool_sx + off_pool);
      const int buf = gi % NB;
      ptx::mbar_wait(&tmem_full[buf], (uint32_t)(gi / NB) & 1u);
      ptx::tc_fence_after();
#pragma unroll
      for (int ci = 0; ci < CPW; ++ci) {
        uint32_t q[3][16];
        const uint32_t ta = tmem_base + (uint32_t)(buf * ACC) + lane_sel + (uint32_t)((c_begin + ci) * 16);
        ptx::tmem_ld16(ta, q[0]);
        ptx::tmem_ld16(ta + N, q[1]);
        ptx::tmem_ld16(ta + 2 * N, q[2]);
        ptx::tmem_ld_wait();
        if (ci == CPW - 1) {                           // last load from this strip's buffer has landed: hand it back before processing
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
        }
        uint32_t h[8];
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + (c_begin + ci) * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 bb = b4[i];
          float f[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(q[0][4 * i + e]), 1);
            const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(q[2][4 * i + e]), 1);
            f[e] = (left + __uint_as_float(q[1][4 * i + e])) + right;
          }
          f[0] += bb.x; f[1] += bb.y; f[2] += bb.z; f[3] += bb.w;
          if (p.relu) { f[0] = fmaxf(f[0], 0.f); f[1] = fmaxf(f[1], 0.f); f[2] = fmaxf(f[2], 0.f); f[3] = fmaxf(f[3], 0.f); }
          __half2 ha = __floats2half2_rn(f[0], f[1]), hb = __floats2half2_rn(f[2], f[3]);
          h[2 * i] = *reinterpret_cast<uint32_t*>(&ha);
          h[2 * i + 1] = *reinterpret_cast<uint32_t*>(&hb);
        }
        if (p.out && valid) ptx::st_global_256(o_full + ci * 16, h);
        if (p.pool_out) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            __half2 a = *reinterpret_cast<__half2*>(&h[i]);
            uint32_t o1 = __shfl_down_sync(0xffffffffu, h[i], 1);
            a = __hmax2(a, *reinterpret_cast<__half2*>(&o1));
            uint32_t cur = *reinterpret_cast<uint32_t*>(&a);
            uint32_t o16 = __shfl_xor_sync(0xffffffffu, cur, 16);
            a = __hmax2(a, *reinterpret_cast<__half2*>(&o16));
            h[i] = *reinterpret_cast<uint32_t*>(&a);
          }
          if (valid && pool_sel) ptx::st_global_256(o_pool + ci * 16, h);
        }
      }
      tx += dx;
      if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
      ty += dy;
      if (ty >= p.tiles_y) { ty -= p.tiles_y; ++tz; }
      tz += dz;
    }
  } else {
    // ===== epilogue: 8 warps per strip; warp -> TMEM lane quarter (warp & 3) and channel half =====
    // Software-pipelined over the flat sequence of (strip, 16-channel chunk) items of this warp: the three TMEM loads of item j + 1 are in
    // flight while item j is shifted / summed / stored from the other register set, and a strip's TMEM buffer is handed back to the MMA warp
    // as soon as the warp's last load from it has completed -- i.e. BEFORE that data is processed.  (First version: load, wait, process,
    // release per strip, all eight warps in lock step: 1450 cycles per strip for ONE chunk per warp, a pure latency chain.)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int xx = lane & 15;                          // input column of this lane inside the 16-wide strip
    const int yr = 2 * quarter + (lane >> 4);          // strip row
    constexpr int CPW = N / 32;                        // chunks of 16 output channels per warp and strip: 1 (N = 32) or 2 (N = 64)
    constexpr int K = S * CPW;                         // items per tile (even: the register set of an item is its parity)
    const int c_begin = half * CPW;
    const uint32_t lane_sel = uint32_t(quarter * 32) << 16;
    const bool x_ok = (xx >= 1) && (xx <= kFoldTX);
    uint32_t r[2][3][16];
    int bufL = 0;                                      // TMEM buffer of the strip whose items are being LOADED
    uint32_t phL = 0;
    auto load3 = [&](uint32_t (&q)[3][16], int ci) {
      const uint32_t ta = tmem_base + (uint32_t)(bufL * ACC) + lane_sel + (uint32_t)((c_begin + ci) * 16);
      ptx::tmem_ld16(ta, q[0]);               // kx = 0 partial sums
      ptx::tmem_ld16(ta + N, q[1]);           // kx = 1
      ptx::tmem_ld16(ta + 2 * N, q[2]);       // kx = 2
    };
    auto release = [&]() {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[bufL]);
      if (++bufL == NB) { bufL = 0; phL ^= 1; }
    };
    int t = blockIdx.x;
    if (t < total_tiles) {
      ptx::mbar_wait(&tmem_full[bufL], phL);
      ptx::tc_fence_after();
      load3(r[0], 0);
      ptx::tmem_ld_wait();
      if (CPW == 1) release();
    }
    for (; t < total_tiles; t += gridDim.x) {
      const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, tz = t / (p.tiles_x * p.tiles_y);
      const int x = tx * kFoldTX - 1 + xx;
      const bool next_tile = t + (int)gridDim.x < total_tiles;
      long long* tre = (p.trace && blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 9) && t / (int)gridDim.x < 64) ? p.trace + (t / gridDim.x) * 8 + (warp == 2 ? 4 : 6) : nullptr;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int s = k / CPW, ci = k % CPW;
        if (tre && k == 0) tre[0] = clock64();
        const bool has_next = (k < K - 1) || next_tile;
        constexpr int dummy = 0; (void)dummy;
        const int nci = (k + 1) % CPW;
        if (has_next) {
          if (nci == 0) { ptx::mbar_wait(&tmem_full[bufL], phL); ptx::tc_fence_after(); }     // the next item opens a new strip
          load3(r[(k + 1) & 1], nci);
        }
        // ---- item k: out[x] = P_0[x - 1] + P_1[x] + P_2[x + 1] (+ bias, ReLU) -> fp16 -> store / 2x2 max-pool ----
        {
          const uint32_t (&q)[3][16] = r[k & 1];
          const int y = ty * 8 * S + 8 * s + yr;
          const bool valid = x_ok && (x < p.W) && (y < p.H);
          const int nbase = (c_begin + ci) * 16;
          uint32_t h[8];
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + nbase);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bb = b4[i];
            float f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(q[0][4 * i + e]), 1);      // P_0 of column x - 1
              const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(q[2][4 * i + e]), 1);   // P_2 of column x + 1
              f[e] = (left + __uint_as_float(q[1][4 * i + e])) + right;
            }
            f[0] += bb.x; f[1] += bb.y; f[2] += bb.z; f[3] += bb.w;
            if (p.relu) { f[0] = fmaxf(f[0], 0.f); f[1] = fmaxf(f[1], 0.f); f[2] = fmaxf(f[2], 0.f); f[3] = fmaxf(f[3], 0.f); }
            __half2 ha = __floats2half2_rn(f[0], f[1]), hb = __floats2half2_rn(f[2], f[3]);
            h[2 * i] = *reinterpret_cast<uint32_t*>(&ha);
            h[2 * i + 1] = *reinterpret_cast<uint32_t*>(&hb);
          }
          if (p.out && valid && nbase < p.n_valid) {
            __half* o = p.out + (long long)tz * p.out_sb + (long long)y * p.out_sy + (long long)x * p.out_sx + nbase;
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
          if (p.pool_out) {
            // 2x2 max-pool on the fp16-rounded values: partners are the next column (lane + 1) and the next row (lane ^ 16)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __half2 a = *reinterpret_cast<__half2*>(&h[i]);
              uint32_t o1 = __shfl_down_sync(0xffffffffu, h[i], 1);
              a = __hmax2(a, *reinterpret_cast<__half2*>(&o1));
              uint32_t cur = *reinterpret_cast<uint32_t*>(&a);
              uint32_t o16 = __shfl_xor_sync(0xffffffffu, cur, 16);
              a = __hmax2(a, *reinterpret_cast<__half2*>(&o16));
              h[i] = *reinterpret_cast<uint32_t*>(&a);
            }
            const bool pool_lane = valid && (xx & 1) && !(lane & 16);     // x even (tiles start at even columns), y even
            if (pool_lane && nbase < p.n_valid) {
              __half* o = p.pool_out + (long long)tz * p.pool_sb + (long long)(y >> 1) * p.pool_sy + (long long)(x >> 1) * p.pool_sx + nbase;
              if (nbase + 16 <= p.n_valid) {
                if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {
                  ptx::st_global_256(o, h);
                } else {
                  *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
                  *reinterpret_cast<uint4*>(o + 8) = make_uint4(h[4], h[5], h[6], h[7]);
                }
              } else {
                unsigned short* os = reinterpret_cast<unsigned short*>(o);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if (nbase + 2 * i < p.n_valid) os[2 * i] = (unsigned short)(h[i] & 0xffffu);
                  if (nbase + 2 * i + 1 < p.n_valid) os[2 * i + 1] = (unsigned short)(h[i] >> 16);
                }
              }
            }
          }
        }
        if (has_next) {
          ptx::tmem_ld_wait();
          if (nci == CPW - 1) release();        // the warp's last load from that strip's buffer has landed: the MMA warp may overwrite it
        }
        if (tre && k == K - 1) tre[1] = clock64();
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, tmem_cols); }
}

}  // namespace airfe
