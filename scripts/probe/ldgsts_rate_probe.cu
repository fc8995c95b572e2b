// Probe 6: how fast can W warps of one CTA per SM stream global memory into shared memory with 16-byte cp.async
// (LDGSTS.128), compared with plain 16-byte loads + st.shared?  Every lane copies 16 bytes; a warp-instruction moves
// 512 contiguous bytes (what the narrow-row halo gather of conv_halo.cu would issue).  Reports B / clk / SM and TB/s
// chip-wide for W = 1, 2, 4, 8 warps and G = 4, 8, 16 instructions in flight per warp between waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: cp.async.cg 16 B, 1: cp.async.cg with zero-fill operand, 2: ld.global.nc.v4 + st.shared.v4
__global__ void __launch_bounds__(256, 1)
probe(const uint4* __restrict__ src, size_t n16_per_cta, int warps, int group, long long* clk_out, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps) return;
  const uint4* base = src + (size_t)blockIdx.x * n16_per_cta;
  const uint32_t ring = 16 * 1024;                       // per-warp ring in shared memory
  const uint32_t s0 = smem_u32(smem) + warp * ring;
  const size_t per_warp = n16_per_cta / warps;
  const uint4* p = base + (size_t)warp * per_warp + lane;
  const size_t iters = per_warp / 32;
  unsigned acc = 0;
  __syncwarp();
  const long long t0 = clock64();
  for (size_t i = 0; i < iters; i += group) {
    for (int g = 0; g < group; ++g) {
      const uint32_t dst = s0 + (uint32_t)(((i + g) * 512 + lane * 16) & (ring - 1));
      if (MODE == 0) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p + (i + g) * 32) : "memory");
      } else if (MODE == 1) {
        const uint32_t sz = ((i + g + lane) & 63) ? 16u : 0u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(p + (i + g) * 32), "r"(sz) : "memory");
      } else {
        const uint4 v = __ldg(p + (i + g) * 32);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        acc += v.x;
      }
    }
    if (MODE != 2) {
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // keep one group (`group` instructions) in flight
    }
  }
  if (MODE != 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
  const long long t1 = clock64();
  if (lane == 0) atomicMax((unsigned long long*)&clk_out[blockIdx.x], (unsigned long long)(t1 - t0));
  if (acc == 0x12345678u) *sink = acc;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <int MODE>
static int run(const char* name, const uint4* src, size_t n16_per_cta, int warps, int group, long long* dclk, unsigned* sink) {
  CK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaMemset(dclk, 0, 148 * 8));
    CK(cudaEventRecord(e0));
    probe<MODE><<<148, 256, 160 * 1024>>>(src, n16_per_cta, warps, group, dclk, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  long long h[148];
  CK(cudaMemcpy(h, dclk, sizeof(h), cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double bytes = (double)n16_per_cta * 16;
  printf("%-34s warps %d group %2d: %6.1f B/clk/SM, %5.1f clk per warp instruction, %5.2f TB/s chip-wide (events)\n", name, warps,
         group, bytes / mx, (double)mx / (n16_per_cta / warps / 32), 148 * bytes / (ms * 1e-3) / 1e12);
  return 0;
}

int main() {
  const size_t n16_per_cta = (size_t)1 << 18;   // 4 MB per CTA, 592 MB in all (streams from HBM)
  uint4* src;
  long long* dclk;
  unsigned* sink;
  CK(cudaMalloc(&src, 148 * n16_per_cta * 16));
  CK(cudaMemset(src, 1, 148 * n16_per_cta * 16));
  CK(cudaMalloc(&dclk, 148 * 8));
  CK(cudaMalloc(&sink, 4));
  for (int warps : {1, 2, 4, 8})
    for (int group : {4, 16}) {
      if (run<0>("cp.async.cg 16 B", src, n16_per_cta, warps, group, dclk, sink)) return 1;
      if (run<1>("cp.async.cg 16 B, zfill operand", src, n16_per_cta, warps, group, dclk, sink)) return 1;
      if (run<2>("ld.global.nc.v4 + st.shared.v4", src, n16_per_cta, warps, group, dclk, sink)) return 1;
    }
  return 0;
}
