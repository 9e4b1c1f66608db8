// Hardware probe (authoring aid, not product): issue rate of tcgen05.mma kind::f16, M=128, K=16, as a function of N, of the A operand
// source (shared memory "SS" vs tensor memory "TS") and of the A shared-memory layout (SWIZZLE_128B canonical, the halo-conv descriptor
// with SBO = 18*128 B, SWIZZLE_64B).  Answers: is a small-N convolution bound by the tensor pipe, or by the A read from shared memory?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_mma_rate tools/probe_mma_rate.cu
//   run:   tools/probe_mma_rate            (one CTA, then one CTA per SM)
#include "../airslam_b200/csrc/ptx.cuh"
#include <cstdio>
#include <vector>
using namespace airfe;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mode 0: SS, A canonical SW128 (SBO 1024), K advance +32 B inside a 64-wide K block, 4 k-steps then next "tap" (row shift)
// mode 1: SS, A halo descriptor (SBO = 18*128), start address walks taps like the conv kernel
// mode 2: SS, A SW64 (64-byte rows, SBO 512), 2 k-steps per block
// mode 3: TS, A in tensor memory (columns 256..263), B from shared memory
// mode 4: SS, same A and B descriptor every time (best-case reuse)
template <int MODE>
__global__ void __launch_bounds__(128, 1) rate(int n, int n_outer, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;                 // 64 KiB region for the A variants
  uint8_t* Bm = smem + 65536;        // one 32 KiB B tile (256 rows x 128 B)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  if (threadIdx.x < 32) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x < 32) {   // whole warp runs the uniform loop, one elected lane issues (see tc_conv3x3.cuh)
    const uint32_t idesc = ptx::make_idesc_f16(128, n, 0);
    const uint32_t sa = ptx::smem_u32(A), sb = ptx::smem_u32(Bm);
    const uint64_t db_c = ptx::smem_desc_base_sw128(1024) + (sb >> 4);
    const uint64_t da_c128 = ptx::smem_desc_base_sw128(1024) + (sa >> 4);
    const uint64_t da_halo = ptx::smem_desc_base_sw128(18 * 128) + (sa >> 4);
    const uint64_t da_sw64 = ptx::smem_desc_base_sw64(512) + (sa >> 4);
    const uint64_t db_sw64 = ptx::smem_desc_base_sw64(512) + (sb >> 4);
    const long long t0 = clock64();
    for (int o = 0; o < n_outer; ++o) {
      if (ptx::elect_one()) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t accum = (o | tap | k) != 0;
          if (MODE == 0) ptx::umma_f16(tb, da_c128 + 2 * k + tap * 64, db_c + 2 * k, idesc, accum);
          else if (MODE == 1) ptx::umma_f16(tb, da_halo + 2 * k + ((tap / 3) * 18 + tap % 3) * 8, db_c + 2 * k, idesc, accum);
          else if (MODE == 2) ptx::umma_f16(tb, da_sw64 + 2 * (k & 1) + ((tap / 3) * 18 + tap % 3) * 4, db_sw64 + 2 * (k & 1), idesc, accum);
          else if (MODE == 3) umma_f16_ts(tb, tb + 256 + 8 * k, db_c + 2 * k, idesc, accum);
          else ptx::umma_f16(tb, da_c128, db_c, idesc, accum);
        }
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::umma_commit(&bar);
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tb, 512); }
}

template <int MODE>
static cudaError_t run(int grid, int n, int n_outer, long long* d) {
  cudaFuncSetAttribute(rate<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  rate<MODE><<<grid, 128, 98304 + 1024>>>(n, n_outer, d);
  return cudaDeviceSynchronize();
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * sms);
  const char* names[] = {"SS sw128 canonical", "SS sw128 halo (SBO 2304)", "SS sw64 halo", "TS (A in TMEM)", "SS same descriptors"};
  const int n_outer = 64, n_mma = n_outer * 36;
  std::vector<long long> h(sms);
  for (int grid : {1, sms}) {
    printf("== grid %d CTA(s), %d MMAs each (M=128, K=16, fp16 -> fp32) ==\n", grid, n_mma);
    for (int mode = 0; mode < 5; ++mode)
      for (int n : {16, 32, 64, 128, 256}) {
        for (int rep = 0; rep < 2; ++rep) {
          cudaError_t e = mode == 0 ? run<0>(grid, n, n_outer, d) : mode == 1 ? run<1>(grid, n, n_outer, d) : mode == 2 ? run<2>(grid, n, n_outer, d)
                        : mode == 3 ? run<3>(grid, n, n_outer, d) : run<4>(grid, n, n_outer, d);
          if (e != cudaSuccess) { printf("mode %d n %d: %s\n", mode, n, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        const double cyc = (double)mx / n_mma;
        printf("%-28s N=%3d : %7.1f cycles/MMA   %6.0f MAC/clk/SM  (floor 128*N/256 = %d)\n", names[mode], n, cyc, 128.0 * n * 16 / cyc, n / 2);
      }
  }
  return 0;
}
