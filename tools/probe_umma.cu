// Hardware probe (authoring aid, not product): how does tcgen05.mma apply SWIZZLE_128B when the A descriptor's start
// address is not 1024-byte aligned and when SBO is not a multiple of 1024?  Decides whether a 3x3 convolution can
// read its nine shifted A operands out of ONE halo tile in shared memory (see DESIGN.md "halo reuse").
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_umma tools/probe_umma.cu
#include "../airslam_b200/csrc/ptx.cuh"
#include <cstdio>
#include <vector>
using namespace airfe;

__global__ void probe(int shift_rows, int sbo, int base_off, float* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __half* A = reinterpret_cast<__half*>(smem);             // 256 rows x 128 B  (32 KiB)
  __half* Bm = reinterpret_cast<__half*>(smem + 32768);    // 64 rows x 128 B   (identity)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 256 * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64, c = k / 8, e = k % 8;
    A[r * 64 + ((c ^ (r & 7)) * 8) + e] = __float2half((float)(r * 8 + c));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64, c = k / 8, e = k % 8;
    Bm[r * 64 + ((c ^ (r & 7)) * 8) + e] = __float2half(r == k ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  if (threadIdx.x < 32) { ptx::tmem_alloc(&slot, 64); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tb = slot;
  if (threadIdx.x == 0) {
    uint32_t idesc = ptx::make_idesc_f16(128, 64, 0);
    uint32_t sa = ptx::smem_u32(A) + shift_rows * 128, sb = ptx::smem_u32(Bm);
    for (int k = 0; k < 4; ++k)
      ptx::umma_f16(tb, ptx::make_smem_desc(sa + k * 32, 16, sbo, 2, base_off), ptx::make_smem_desc(sb + k * 32, 16, 1024, 2), idesc, k != 0);
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int c = 0; c < 64; c += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c, r);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tb, 64);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int cfgs[][3] = {{0, 1024, 0}, {1, 1024, 0}, {1, 1024, 1}, {2, 1024, 0}, {2, 1024, 2}, {0, 1280, 0}, {1, 1280, 0}, {1, 1280, 1},
                   {0, 1152, 0}, {3, 2304, 0}, {3, 2304, 3}, {8, 1024, 0}, {0, 2048, 0}, {1, 2048, 0}, {1, 2048, 1}};
  std::vector<float> h(128 * 64);
  for (auto& c : cfgs) {
    cudaMemset(d, 0, 128 * 64 * 4);
    probe<<<1, 128, 49152 + 1024>>>(c[0], c[1], c[2], d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cfg shift=%d sbo=%d base_off=%d : %s\n", c[0], c[1], c[2], cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(h.data(), d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    // per M row: source row = value/8 at chunk 0, and chunk consistency across the 8 chunks
    for (int m = 0; m < 128; ++m) {
      int row0 = (int)h[m * 64] / 8;
      bool ok = true;
      for (int c8 = 0; c8 < 8; ++c8)
        for (int e = 0; e < 8; ++e)
          if ((int)h[m * 64 + c8 * 8 + e] != row0 * 8 + c8) ok = false;
      printf("%d%s ", row0, ok ? "" : "!");
      if (m % 16 == 15) printf("\n");
    }
    // detail for bad rows (first 3)
    int shown = 0;
    for (int m = 0; m < 128 && shown < 3; ++m) {
      int row0 = (int)h[m * 64] / 8;
      bool ok = true;
      for (int c8 = 0; c8 < 8; ++c8) if ((int)h[m * 64 + c8 * 8] != row0 * 8 + c8) ok = false;
      if (!ok) { printf("  row m=%d chunks:", m); for (int c8 = 0; c8 < 8; ++c8) printf(" %d", (int)h[m * 64 + c8 * 8]); printf("\n"); ++shown; }
    }
  }
  return 0;
}
