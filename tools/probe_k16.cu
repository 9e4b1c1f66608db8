// Hardware probe (authoring aid, not product; prepared in round 1 for the round-2 conv1a-on-tensor-core work, NOT YET RUN):
// which shared-memory layout / descriptor does tcgen05.mma accept for a K = 16 fp16 A operand (one K step, 32-byte rows)?
// conv1a (1 -> 64 channels, 3x3) as an implicit GEMM has K = 9 taps padded to 16: the im2col tile is [128 pixels x 16] fp16 = 32 bytes
// per pixel.  Candidates tried here, each validated against a host product D = A . B^T with B = [64 x 16] in the same layout:
//   0  SWIZZLE_NONE, K-major "interleaved" canonical layout: core matrix = 8 rows x 16 bytes contiguous (128 B);
//      element (r, k) at (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2          with LBO = 128, SBO = 256
//   1  same with LBO = 8*rows*16 (all k=0 core matrices first, then all k=1): LBO = rows*16, SBO = 128
//   2  SWIZZLE_32B, 32-byte rows: element (r, k) at r*32 + ((k/8) ^ ((r >> 2) & 1))*16 + (k%8)*2, SBO = 256 (8 rows), layout type 6
//   3  SWIZZLE_32B with the XOR taken from address bit 7 ( (r >> 2) & 1 is address bit 7 for 32-byte rows: same thing written on the
//      absolute address, to confirm the swizzle is address-based like the 128B mode: tile placed at a 128-byte (not 256) offset)
// Output: per candidate the number of mismatching elements of the 128 x 64 result (0 = layout understood).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_k16 tools/probe_k16.cu
#include "../airslam_b200/csrc/ptx.cuh"
#include <cstdio>
#include <cstring>
#include <vector>
using namespace airfe;

struct Cand { int layout_type; int lbo, sbo_a, sbo_b; int base_off; };

__global__ void probe(const __half* a_img, const __half* b_img, int a_bytes, int b_bytes, int layout_type, int lbo, int sbo_a, int sbo_b, int a_off,
                      float* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem + a_off;              // a_off lets a candidate start the tile off a 256-byte boundary
  uint8_t* Bm = smem + 16384;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < a_bytes; i += blockDim.x) A[i] = reinterpret_cast<const uint8_t*>(a_img)[i];
  for (int i = threadIdx.x; i < b_bytes; i += blockDim.x) Bm[i] = reinterpret_cast<const uint8_t*>(b_img)[i];
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  if (threadIdx.x < 32) { ptx::tmem_alloc(&slot, 64); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x < 32) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_f16(128, 64, 0);
      ptx::umma_f16(tb, ptx::make_smem_desc(ptx::smem_u32(A), lbo, sbo_a, layout_type), ptx::make_smem_desc(ptx::smem_u32(Bm), lbo, sbo_b, layout_type), idesc, 0);
      ptx::umma_commit(&bar);
    }
    __syncwarp();
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int c = 0; c < 64; c += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c, r);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tb, 64); }
}

static size_t place(int cand, int rows, int r, int k, int a_off_bits) {
  switch (cand) {
    case 0: return (size_t)(r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
    case 1: return (size_t)(k / 8) * rows * 16 + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
    case 2: return (size_t)r * 32 + (((k / 8) ^ ((r >> 2) & 1)) * 16) + (k % 8) * 2;
    default: {   // address-based XOR: chunk ^= bit 7 of the absolute (tile-relative + a_off) address
      const size_t lin = (size_t)r * 32 + a_off_bits;
      return (size_t)r * 32 + (((k / 8) ^ ((lin >> 7) & 1)) * 16) + (k % 8) * 2;
    }
  }
}

int main() {
  std::vector<float> a(128 * 16), b(64 * 16), ref(128 * 64), h(128 * 64);
  for (int r = 0; r < 128; ++r) for (int k = 0; k < 16; ++k) a[r * 16 + k] = (float)((r * 7 + k * 3) % 13 - 6);
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 16; ++k) b[n * 16 + k] = (float)((n * 5 + k * 11) % 9 - 4);
  for (int r = 0; r < 128; ++r) for (int n = 0; n < 64; ++n) { float s = 0; for (int k = 0; k < 16; ++k) s += a[r * 16 + k] * b[n * 16 + k]; ref[r * 64 + n] = s; }
  __half *da, *db; float* dout;
  cudaMalloc(&da, 16384); cudaMalloc(&db, 16384); cudaMalloc(&dout, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  const char* names[] = {"no swizzle, LBO 128 / SBO 256", "no swizzle, LBO rows*16 / SBO 128", "SWIZZLE_32B row-index XOR", "SWIZZLE_32B address XOR, tile at +128 B"};
  for (int cand = 0; cand < 4; ++cand) {
    const int a_off = cand == 3 ? 128 : 0;
    std::vector<__half> ai(8192, __float2half(0.f)), bi(8192, __float2half(0.f));
    for (int r = 0; r < 128; ++r) for (int k = 0; k < 16; ++k) ai[place(cand, 128, r, k, a_off) / 2] = __float2half(a[r * 16 + k]);
    for (int n = 0; n < 64; ++n) for (int k = 0; k < 16; ++k) bi[place(cand == 3 ? 2 : cand, 64, n, k, 0) / 2] = __float2half(b[n * 16 + k]);
    cudaMemcpy(da, ai.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 128 * 64 * 4);
    int lt = 0, lbo = 128, sbo_a = 256, sbo_b = 256;
    if (cand == 1) { lbo = 128 * 16; sbo_a = 128; sbo_b = 128; }
    if (cand == 1) { /* B has 64 rows: its LBO differs; the descriptor field is shared in this probe, so B is laid out with rows = 128 too */ }
    if (cand >= 2) { lt = 6; lbo = 16; sbo_a = 256; sbo_b = 256; }
    if (cand == 1) for (int n = 0; n < 64; ++n) for (int k = 0; k < 16; ++k) bi[place(1, 128, n, k, 0) / 2] = __float2half(b[n * 16 + k]);
    if (cand == 1) cudaMemcpy(db, bi.data(), 16384, cudaMemcpyHostToDevice);
    probe<<<1, 128, 33 * 1024>>>(da, db, 8192, 8192, lt, lbo, sbo_a, sbo_b, a_off, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cand %d (%s): %s\n", cand, names[cand], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 128 * 64; ++i) bad += (h[i] != ref[i]);
    printf("cand %d  %-45s : %d / %d mismatching elements%s\n", cand, names[cand], bad, 128 * 64, bad ? "" : "   <-- layout understood");
  }
  return 0;
}
