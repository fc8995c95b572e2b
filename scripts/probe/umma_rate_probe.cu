// Probe 5: back-to-back tcgen05.mma rate on B200 as a function of
//   * cta_group (1: M = 128 on one SM; 2: M = 256 on an SM pair, B split across the pair),
//   * N (64 .. 256), operand majorness (K-major as fprop / dgrad, MN-major as wgrad),
//   * where A comes from (shared memory or TMEM),
//   * the issue pattern (GEMM ring / wgrad: several accumulators share B / fprop: taps share the halo tile).
// Operand CONTENT is irrelevant (timing only).  One elected thread issues `iters` rounds of the pattern and
// waits for one commit; clk / MMA = elapsed / count.  Run with grid = 1 cluster and with the whole chip.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../mmr_semantic-segmentation_v1_b200/csrc/ptx.cuh"
using namespace mmr;

struct Cfg {
  int cg;           // cta_group
  int N;            // MMA N (whole pair for cg = 2)
  int a_tmem;       // A operand from TMEM
  uint32_t a_hi, b_hi;        // descriptor high words (SBO, version, swizzle)
  uint32_t a_lo0, b_lo0;      // descriptor low words without the start address (LBO << 16)
  int outer, mid, inner;      // loop trip counts of one round
  int ao, am, ak;             // A start-address steps (16-byte units) per outer / mid / inner index
  int bo, bm, bk;             // B
  int dm;                     // accumulator column step per mid index
  int iters;
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma1(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma2(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma1_ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 db, {%2, %3};\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}\n" ::"r"(d), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 db, {%2, %3};\n"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n}\n" ::"r"(d), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int CG>
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ Cfg c, long long* clk_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 192 * 1024);
  uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x)
    ((uint32_t*)smem)[i] = ((i * 2654435761u) & 0x80008000u) | 0x3c003c00u | ((i * 40503u) & 0x007f007fu);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
  if (warp == 0) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(tptr, 512);
      tmem_relinquish();
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, *tptr, 0);
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(128 * CG, c.N, (c.a_lo0 >> 31) & 1, (c.b_lo0 >> 31) & 1);
    const uint32_t a0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | (c.a_lo0 & 0x3FFF0000u);
    const uint32_t b0 = (((smem_u32(smem) + 128 * 1024) >> 4) & 0x3FFFu) | (c.b_lo0 & 0x3FFF0000u);
    const uint32_t a_t = tb + 448;  // A-in-TMEM: the last 64 columns (K = 16 bf16 = 8 columns per MMA)
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < c.iters; ++it)
        for (int o = 0; o < c.outer; ++o)
          for (int k = 0; k < c.inner; ++k)
#pragma unroll 1
            for (int m = 0; m < c.mid; ++m) {
              const uint32_t alo = a0 + (uint32_t)(o * c.ao + m * c.am + k * c.ak);
              const uint32_t blo = b0 + (uint32_t)(o * c.bo + m * c.bm + k * c.bk);
              const uint32_t d = tb + (uint32_t)(m * c.dm);
              if (CG == 2) {
                if (c.a_tmem) mma2_ts(d, a_t + (uint32_t)(k & 3) * 8, blo, c.b_hi, idesc);
                else mma2(d, alo, c.a_hi, blo, c.b_hi, idesc);
              } else {
                if (c.a_tmem) mma1_ts(d, a_t + (uint32_t)(k & 3) * 8, blo, c.b_hi, idesc);
                else mma1(d, alo, c.a_hi, blo, c.b_hi, idesc);
              }
            }
      if (CG == 2) commit2(bar); else umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (elect_one()) clk_out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  if (warp == 0) {
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
    else tmem_dealloc(tb, 512);
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

static uint32_t hi_word(uint32_t sbo_bytes, uint32_t swz) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (swz << 29); }


// Lean variant: everything but the stage base is a compile-time constant, loops fully unrolled, so the
// issuing thread spends ~4 instructions per MMA (the generic kernel above needs ~20 and is issue-bound
// below ~80 clk / MMA).  One round = STAGES x KSTEPS x NACC MMAs; accumulator m at column m * N.
template <int N, int NACC, int KSTEPS, int STAGES, int AM, int AK, int AO, int BK, int BO, int AMN, int BMN>
__global__ void __launch_bounds__(128, 1)
probe_lean(uint32_t a_hi, uint32_t b_hi, uint32_t a_lbo, uint32_t b_lbo, int iters, long long* clk_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 192 * 1024);
  uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x)
    ((uint32_t*)smem)[i] = ((i * 2654435761u) & 0x80008000u) | 0x3c003c00u | ((i * 40503u) & 0x007f007fu);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, *tptr, 0);
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N, AMN, BMN);
    const uint32_t a0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | (a_lbo << 16);
    const uint32_t b0 = (((smem_u32(smem) + 128 * 1024) >> 4) & 0x3FFFu) | (b_lbo << 16);
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
#pragma unroll 1
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k)
#pragma unroll
            for (int m = 0; m < NACC; ++m)
              mma1(tb + m * N, a0 + s * AO + k * AK + m * AM, a_hi, b0 + s * BO + k * BK, b_hi, idesc);
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (elect_one()) clk_out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int NACC, int KSTEPS, int STAGES, int AM, int AK, int AO, int BK, int BO, int AMN, int BMN>
static int run_lean(const char* name, uint32_t a_hi, uint32_t b_hi, uint32_t a_lbo, uint32_t b_lbo, int grid, long long* dclk) {
  auto kern = probe_lean<N, NACC, KSTEPS, STAGES, AM, AK, AO, BK, BO, AMN, BMN>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int iters = 256;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    kern<<<grid, 128, 200 * 1024>>>(a_hi, b_hi, a_lbo, b_lbo, iters, dclk);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  long long h[148];
  CK(cudaMemcpy(h, dclk, 148 * 8, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double count = (double)iters * STAGES * KSTEPS * NACC;
  const double per = (double)mx / count;
  printf("lean %-44s grid %3d N %3d acc %d: %6.1f clk/MMA (floor %5.1f) %3.0f%%  [%.0f TFLOP/s by events]\n", name, grid, N, NACC, per, N / 2.0,
         100.0 * (N / 2.0) / per, (double)grid * count * 128.0 * N * 32 / (ms * 1e-3) / 1e12);
  return 0;
}

static int run(const char* name, Cfg c, int grid, long long* dclk) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = c.cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = c.cg == 2 ? 1 : 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(dclk, 0, 148 * 8));
    CK(cudaEventRecord(e0));
    if (c.cg == 2) CK(cudaLaunchKernelEx(&cfg, probe<2>, c, dclk)); else CK(cudaLaunchKernelEx(&cfg, probe<1>, c, dclk));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  long long h[148];
  CK(cudaMemcpy(h, dclk, 148 * 8, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double count = (double)c.iters * c.outer * c.mid * c.inner;
  const double per = (double)mx / count;
  const double floor_clk = c.N / 2.0;  // 128 x N x 16 MACs per SM at 4096 MAC / clk / SM
  const double tflops = (double)grid * count * 128.0 * c.N * 16 * 2 / (ms * 1e-3) / 1e12;
  printf("%-34s grid %3d cg %d N %3d %s: %6.1f clk/MMA (floor %5.1f) %3.0f%%   [%.0f TFLOP/s by events incl. launch]\n", name, grid, c.cg,
         c.N, c.a_tmem ? "A=TMEM" : "A=smem", per, floor_clk, 100.0 * floor_clk / per, tflops);
  return 0;
}

int main() {
  long long* dclk;
  CK(cudaMalloc(&dclk, 148 * 8));
  CK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const uint32_t MN = 0x80000000u;  // flag in *_lo0: MN-major

  // ---- lean patterns (cta_group::1, A and B in shared memory)
  for (int grid : {1, 148}) {
    const uint32_t hK = hi_word(1024, 2), hH = hi_word(10 * 128, 2), hZ = hi_word(8 * 128, 2), hZ10 = hi_word(10 * 128, 2);
    // K-major fprop-like: 4 accumulators (M tiles 16 halo rows apart), 4 K steps of 32 B, 2 "taps" as stages
    if (run_lean<64, 4, 4, 2, ((16 * 10 * 128) / 16), 2, 8, 2, 0, 0, 0>("fprop K-major 4 tiles", hH, hK, 1, 1, grid, dclk)) return 1;
    if (run_lean<128, 4, 4, 2, ((16 * 10 * 128) / 16), 2, 8, 2, 0, 0, 0>("fprop K-major 4 tiles", hH, hK, 1, 1, grid, dclk)) return 1;
    if (run_lean<192, 2, 4, 2, ((16 * 10 * 128) / 16), 2, 8, 2, 0, 0, 0>("fprop K-major 2 tiles", hH, hK, 1, 1, grid, dclk)) return 1;
    if (run_lean<256, 2, 4, 2, ((16 * 10 * 128) / 16), 2, 8, 2, 0, 0, 0>("fprop K-major 2 tiles", hH, hK, 1, 1, grid, dclk)) return 1;
    if (run_lean<256, 1, 4, 2, ((16 * 10 * 128) / 16), 2, 8, 2, 0, 0, 0>("K-major one accumulator (dependent chain)", hH, hK, 1, 1, grid, dclk)) return 1;
    if (run_lean<64, 2, 4, 2, ((16 * 10 * 128) / 16), 2, 8, 2, 0, 0, 0>("fprop K-major 2 tiles", hH, hK, 1, 1, grid, dclk)) return 1;
    if (run_lean<64, 8, 4, 2, ((16 * 10 * 128) / 16), 2, 8, 2, 0, 0, 0>("fprop K-major 8 tiles", hH, hK, 1, 1, grid, dclk)) return 1;
    // MN-major wgrad, today: N = 64 (co), 5 accumulators = tap pairs (A atoms 1 pixel apart), K step = 2 halo rows
    if (run_lean<64, 5, 8, 2, ((2 * 128) / 16), ((2 * 10 * 128) / 16), ((32 * 1024) / 16), ((2 * 8 * 128) / 16), ((32 * 1024) / 16), 1, 1>(
            "wgrad today: taps on A", hH, hZ, ((1 * 128) / 16), ((16 * 1024) / 16), grid, dclk)) return 1;
    // wgrad, kx on B: N = 192 = 3 shifted dz atoms (LBO 1 pixel), 2 accumulators = (ky0, ky1) and (ky2, -)
    if (run_lean<192, 2, 8, 2, ((2 * 10 * 128) / 16), ((2 * 10 * 128) / 16), ((32 * 1024) / 16), ((2 * 10 * 128) / 16), ((32 * 1024) / 16), 1, 1>(
            "wgrad kx on B: N = 3 x 64", hH, hZ10, ((10 * 128) / 16), ((1 * 128) / 16), grid, dclk)) return 1;
    // the same with N = 128 (kx 0, 1) and N = 64 (kx 2) is covered by the lines above; N = 256 = 2 kx x 128 co
    if (run_lean<256, 2, 8, 2, ((2 * 10 * 128) / 16), ((2 * 10 * 128) / 16), ((32 * 1024) / 16), ((2 * 10 * 128) / 16), ((32 * 1024) / 16), 1, 1>(
            "wgrad N = 256", hH, hZ10, ((10 * 128) / 16), ((1 * 128) / 16), grid, dclk)) return 1;
    if (run_lean<128, 3, 8, 2, ((2 * 10 * 128) / 16), ((2 * 10 * 128) / 16), ((32 * 1024) / 16), ((2 * 10 * 128) / 16), ((32 * 1024) / 16), 1, 1>(
            "wgrad N = 128, 3 accumulators", hH, hZ10, ((10 * 128) / 16), ((1 * 128) / 16), grid, dclk)) return 1;
    if (run_lean<96, 5, 8, 2, ((2 * 128) / 16), ((2 * 10 * 128) / 16), ((32 * 1024) / 16), ((2 * 8 * 128) / 16), ((32 * 1024) / 16), 1, 1>(
            "wgrad N = 96, 5 accumulators", hH, hZ, ((1 * 128) / 16), ((16 * 1024) / 16), grid, dclk)) return 1;
  }
  if (getenv("PROBE_GENERIC"))
  for (int grid : {2, 148}) {
    for (int cg : {1, 2}) {
      // ---- GEMM ring, K-major SWIZZLE_128B: 4 stages x 4 K steps, one accumulator
      for (int N : {64, 128, 192, 256}) {
        Cfg c = {};
        c.cg = cg; c.N = N; c.a_hi = hi_word(1024, 2); c.b_hi = hi_word(1024, 2); c.a_lo0 = 1u << 16; c.b_lo0 = 1u << 16;
        c.outer = 4; c.mid = 1; c.inner = 4; c.ao = ((16 * 1024) / 16); c.ak = 2; c.bo = ((N / cg) * 128) >> 4; c.bk = 2;
        if (c.bo * 4 * 16 > 64 * 1024) c.bo = 0;
        c.iters = 512;
        if (run("gemm ring K-major", c, grid, dclk)) return 1;
        c.a_tmem = 1;
        if (run("gemm ring K-major", c, grid, dclk)) return 1;
      }
      // ---- fprop-like: 9 taps (shifted views of one halo tile, pitch 10) x T M-tiles x 4 K steps; accumulator per M tile
      for (int N : {64, 128, 192, 256}) {
        const int T = N <= 128 ? 4 : 2;
        Cfg c = {};
        c.cg = cg; c.N = N; c.a_hi = hi_word(10 * 128, 2); c.b_hi = hi_word(1024, 2); c.a_lo0 = 1u << 16; c.b_lo0 = 1u << 16;
        c.outer = 9; c.mid = T; c.inner = 4; c.ao = ((128) / 16); c.am = ((16 * 10 * 128) / 16); c.ak = 2; c.bo = 0; c.bm = 0; c.bk = 2;
        c.dm = N; c.iters = 64;
        if (run("fprop taps x tiles (mid = tiles)", c, grid, dclk)) return 1;
      }
      // ---- wgrad-like: MN-major both, 8 K steps x A accumulators sharing B
      for (int N : {64, 96, 128}) {
        const int A = N * 5 <= 512 ? 5 : (N * 4 <= 512 ? 4 : 3);
        Cfg c = {};
        c.cg = cg; c.N = N;
        c.a_hi = hi_word(10 * 128, 2); c.a_lo0 = MN | ((((1 * 128) / 16)) << 16);
        c.b_hi = hi_word(8 * 128, 2);  c.b_lo0 = MN | ((((16 * 1024) / 16)) << 16);
        c.outer = 2; c.mid = A; c.inner = 8; c.ao = ((32 * 1024) / 16); c.am = ((2 * 128) / 16); c.ak = ((2 * 10 * 128) / 16);
        c.bo = ((32 * 1024) / 16); c.bm = 0; c.bk = ((2 * 8 * 128) / 16); c.dm = N; c.iters = 128;
        if (run("wgrad MN-major (mid = accumulators)", c, grid, dclk)) return 1;
      }
    }
  }
  return 0;
}
