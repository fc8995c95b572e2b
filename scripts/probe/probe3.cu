// Probe 3: lean tcgen05.mma issue loop (uniform operands, precomputed descriptors) to find the
// real SS-mode rate of M=128 x N x K=16 bf16 MMAs whose A operand is a shifted halo view.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../mmr_semantic-segmentation_v1_b200/csrc/ptx.cuh"
using namespace mmr;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::
               "r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
}

template <int T>
__global__ void __launch_bounds__(128, 1)
probeD(int N, int iters, int pitch_rows, long long* clk_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 192 * 1024);
  uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 0xff);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, *tptr, 0);
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 96 * 1024;
    const uint64_t hiA = ((uint64_t)((pitch_rows * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | (1u << 16);
    const uint64_t hiB = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | (1u << 16);
    const uint64_t dA0 = hiA | (a0 >> 4), dB0 = hiB | (b0 >> 4);
    const uint32_t tile_step = (uint32_t)(16 * pitch_rows * 128) >> 4;   // next M-tile: 16 halo rows down
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const uint32_t tapoff = (uint32_t)(((t / 3) * pitch_rows + (t % 3)) * 128) >> 4;
#pragma unroll
        for (int i = 0; i < T; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (elect_one()) umma_acc(tb + i * N, dA0 + tapoff + i * tile_step + 2 * k, dB0 + (t & 1) * ((N * 128) >> 4) + 2 * k, idesc);
          }
        }
      }
    }
    if (elect_one()) umma_commit(bar);
    mbar_wait(bar, 0);
    if (threadIdx.x == 32) clk_out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tb, 512);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
template <int T> int run(int N, int pitch, int grid, long long* dclk) {
  CK(cudaFuncSetAttribute(probeD<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int iters = 200;
  for (int rep = 0; rep < 2; ++rep) { probeD<T><<<grid, 128, 200 * 1024>>>(N, iters, pitch, dclk); CK(cudaDeviceSynchronize()); }
  long long h[148]; CK(cudaMemcpy(h, dclk, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / (iters * 36.0 * T);
  printf("D: grid %3d T %d N %3d pitch %2d: %.1f clk/MMA (floor %d) -> %.0f%% of tensor peak\n", grid, T, N, pitch, per, N / 2, 100.0 * (N / 2) / per);
  return 0;
}
int main() {
  long long* dclk; CK(cudaMalloc(&dclk, 148 * 8));
  for (int grid : {1, 148}) {
    for (int N : {16, 32, 64, 128}) { if (run<4>(N, 18, grid, dclk)) return 1; }
    for (int N : {64, 128, 256}) { if (run<2>(N, 18, grid, dclk)) return 1; }
    for (int N : {64, 128, 256}) { if (run<1>(N, 10, grid, dclk)) return 1; }
    if (run<4>(64, 34, grid, dclk)) return 1;
  }
  return 0;
}
