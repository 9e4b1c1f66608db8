// Hardware probe (authoring aid, not product): throughput of tcgen05.ld (TMEM -> registers), shape 32x32b, as a function of the vector
// width (x16 / x32 / x64 / x128 columns per instruction), of the number of warps reading concurrently (1, 4 = one per lane quarter, 8 = two
// per lane quarter) and of load batching (one load then wait, or two loads in flight per wait).  Answers: what bounds an epilogue that reads
// a 128 x N fp32 accumulator -- bytes, or instructions?  (The kx-folded conv epilogue reads 3x the columns of the nine-tap kernel and was
// measured at ~820 cycles per 48 KB: 58 B/clk.)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_ldtm tools/probe_ldtm.cu
//   run:   tools/probe_ldtm
#include "../airslam_b200/csrc/ptx.cuh"
#include <cstdio>
using namespace airfe;

template <int X>
__device__ __forceinline__ void ldtm(uint32_t taddr, uint32_t& sink) {
  if constexpr (X == 16) {
    uint32_t r[16];
    ptx::tmem_ld16(taddr, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) sink ^= r[i];
  } else if constexpr (X == 32) {
    uint32_t r[32];
    ptx::tmem_ld32(taddr, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= r[i];
  } else {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]),
          "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]),
          "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 64; ++i) sink ^= r[i];
  }
}

// BATCH = 2: two loads in flight before one wait (x16 / x32 only)
template <int X, int BATCH>
__global__ void __launch_bounds__(256, 1) probe(int nwarps, int iters, long long* out, uint32_t* sink_out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      if constexpr (BATCH == 1) {
        ldtm<X>(tb + ((it * X) & 255), sink);
      } else if constexpr (X == 16) {
        uint32_t a[16], b[16], c[16];
        ptx::tmem_ld16(tb + ((it * 48) & 255), a);
        ptx::tmem_ld16(tb + ((it * 48 + 16) & 255), b);
        ptx::tmem_ld16(tb + ((it * 48 + 32) & 255), c);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) sink ^= a[i] ^ b[i] ^ c[i];
      } else {
        uint32_t a[32], b[32];
        ptx::tmem_ld32(tb + ((it * 64) & 255), a);
        ptx::tmem_ld32(tb + ((it * 64 + 32) & 255), b);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) sink ^= a[i] ^ b[i];
      }
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
  sink_out[threadIdx.x] = sink;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(slot, 512); }
}

template <int X, int BATCH>
void run(const char* name, int cols_per_iter) {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 8 * sizeof(long long)); cudaMalloc(&s, 256 * 4);
  const int iters = 2000;
  for (int nw : {1, 4, 8}) {
    probe<X, BATCH><<<1, 256>>>(nw, iters, d, s);
    cudaDeviceSynchronize();
    probe<X, BATCH><<<1, 256>>>(nw, iters, d, s);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
    long long h[8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = (double)nw * iters * cols_per_iter * 32 * 4;
    printf("%-22s warps %d : %7.1f cycles per iteration per warp, %7.1f B/clk per SM\n", name, nw, (double)mx / iters, bytes / mx);
  }
  cudaFree(d); cudaFree(s);
}

int main() {
  run<16, 1>("x16, wait each", 16);
  run<32, 1>("x32, wait each", 32);
  run<64, 1>("x64, wait each", 64);
  run<16, 2>("3 x x16 per wait", 48);
  run<32, 2>("2 x x32 per wait", 64);
  return 0;
}
