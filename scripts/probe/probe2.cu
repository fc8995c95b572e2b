// Probe 2 (design inputs for the halo-tile conv kernel); run on a B200 through gpurun.
//  A. Does cuTensorMapEncodeTiled accept byte stride 0 (replicating dims)?  If so, a 5-D box
//     (C, dupx, Wl, dupy, rows) lands a nearest-x2 upsampled tile in shared memory.
//  B. Shifted / strided K-major operand views for 64-byte and 32-byte rows (SWIZZLE_64B/32B).
//  C. L2 -> shared-memory TMA bandwidth per SM with all SMs loading halo-sized boxes.
//  D. Back-to-back tcgen05.mma rate for M=128, N in {16..256}, operands in shared memory.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../mmr_semantic-segmentation_v1_b200/csrc/ptx.cuh"
using namespace mmr;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc;

__device__ __forceinline__ void tma_load_5d(void* dst, const void* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------- A
__global__ void probeA(const __grid_constant__ CUtensorMap map, int x0, int row0, int bytes, uint4* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 32 * 1024);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, bytes);
    tma_load_5d(smem, &map, bar, 0, 0, x0, 0, row0);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) out[i] = ((uint4*)smem)[i];
}

// ---------------------------------------------------------------- B
__global__ void __launch_bounds__(128, 1)
probeB(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int rowsA, int rowbytes,
       int shift, int pitch, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 48 * 1024;
  uint64_t* bar = (uint64_t*)(smem + 64 * 1024);
  uint64_t* bar2 = bar + 1;
  uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = rowbytes / 2;  // = N
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tptr, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, rowsA * rowbytes + K * rowbytes);
    tma_load_2d(sA, &mapA, bar, 0, 0);
    tma_load_2d(sB, &mapB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, K, 0, 0);
    const uint32_t swz = swizzle_code(rowbytes);
    const uint32_t a0 = smem_u32(sA) + shift * rowbytes;
    const uint32_t b0 = smem_u32(sB);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t da = make_smem_desc(a0 + k * 32, 16, pitch * rowbytes, swz);
      const uint64_t db = make_smem_desc(b0 + k * 32, 16, 8 * rowbytes, swz);
      umma_bf16(tb, da, db, idesc, k != 0);
    }
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < K; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tb, 64);
}

// ---------------------------------------------------------------- C
__global__ void __launch_bounds__(64, 1)
probeC(const __grid_constant__ CUtensorMap map, int stages, int stage_bytes, int box_bytes, int iters, int tiles_x,
       int tiles_y, int n_img, int cchunks, int bw, int bh, long long* clk_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + 8;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int total = tiles_x * tiles_y * n_img * cchunks;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    int t = (blockIdx.x * 37) % total;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&empty[stage], phase ^ 1);
      int q = t;
      const int cc = q % cchunks; q /= cchunks;
      const int tx = q % tiles_x; q /= tiles_x;
      const int ty = q % tiles_y; q /= tiles_y;
      mbar_arrive_expect_tx(&full[stage], box_bytes);
      tma_load_4d(smem + (size_t)stage * stage_bytes, &map, &full[stage], cc * 64, tx * bw - 1, ty * bh - 1, q);
      t += gridDim.x; if (t >= total) t -= total;
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    clk_out[blockIdx.x] = clock64() - t0;
  }
}

// ---------------------------------------------------------------- D
__global__ void __launch_bounds__(128, 1)
probeD(int N, int n_mma, int pitch_rows, int nviews, int n_acc, long long* clk_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;               // up to 96 KB halo
  uint8_t* sB = smem + 96 * 1024;   // up to 9 x 32 KB? keep 64 KB
  uint64_t* bar = (uint64_t*)(smem + 192 * 1024);
  uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 0xff);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    long long t0 = clock64();
    int v = 0, acc = 0;
    for (int i = 0; i < n_mma; i += 4) {
      const uint32_t av = a0 + (uint32_t)((v / 3) * pitch_rows + (v % 3)) * 128;
      const uint32_t bv = b0 + (uint32_t)(v % 2) * (uint32_t)(N * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = make_smem_desc(av + k * 32, 16, pitch_rows * 128, 2);
        const uint64_t db = make_smem_desc(bv + k * 32, 16, 1024, 2);
        umma_bf16(tb + acc * N, da, db, idesc, 1);
      }
      if (++acc == n_acc) { acc = 0; if (++v == nviews) v = 0; }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    clk_out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tb, 512);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  enc = (EncodeTiledFn)fp;
  cuuint32_t es5[5] = {1, 1, 1, 1, 1};

  // ---------------- A: zero-stride replication
  {
    const int N = 2, Hl = 6, Wl = 6, C = 64;
    std::vector<__nv_bfloat16> h((size_t)N * Hl * Wl * C);
    for (int n = 0; n < N; ++n) for (int y = 0; y < Hl; ++y) for (int x = 0; x < Wl; ++x) for (int c = 0; c < C; ++c)
      h[(((size_t)n * Hl + y) * Wl + x) * C + c] = __float2bfloat16((float)(1 + c + 64 * (x + 8 * (y + 8 * n)) % 4096));
    __nv_bfloat16* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap m;
    cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)Wl, 2, (cuuint64_t)(N * Hl)};
    cuuint64_t str[4] = {0, (cuuint64_t)C * 2, 0, (cuuint64_t)Wl * C * 2};
    cuuint32_t box[5] = {64, 2, 4, 2, 3};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, str, box, es5, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("A: encode with zero strides -> CUresult %d\n", (int)r);
    if (r == CUDA_SUCCESS) {
      const int px = 8 * 6, bytes = px * 128;
      uint4* dout; CK(cudaMalloc(&dout, bytes));
      CK(cudaFuncSetAttribute(probeA, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
      for (int x0 : {0, -1, 3}) for (int row0 : {1, 4}) {
        probeA<<<1, 128, 40 * 1024>>>(m, x0, row0, bytes, dout);
        CK(cudaDeviceSynchronize());
        std::vector<__nv_bfloat16> o(bytes / 2);
        CK(cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int p = 0; p < px; ++p) {
          const int ur = p / 8, uc = p % 8;          // upsampled row / col inside the box
          const int ly = row0 + ur / 2, lx = x0 + uc / 2;
          for (int c = 0; c < 64; ++c) {
            const int j = c / 8, phys = p * 64 + ((j ^ (p & 7)) * 8) + (c % 8);
            float want = 0.f;
            if (lx >= 0 && lx < Wl && ly >= 0 && ly < N * Hl) want = __bfloat162float(h[((size_t)ly * Wl + lx) * C + c]);
            if (__bfloat162float(o[phys]) != want) ++bad;
          }
        }
        printf("A: x0 %2d row0 %d: %s (bad %d)\n", x0, row0, bad ? "MISMATCH" : "ok", bad);
      }
    }
  }

  // ---------------- B: 64-byte and 32-byte rows
  for (int rowbytes : {64, 32}) {
    const int rowsA = 512, K = rowbytes / 2;
    std::vector<__nv_bfloat16> hA((size_t)rowsA * K), hB((size_t)K * K);
    for (int r = 0; r < rowsA; ++r) for (int c = 0; c < K; ++c) hA[(size_t)r * K + c] = __float2bfloat16((float)(r + (c % 7) * 512));
    for (int r = 0; r < K; ++r) for (int c = 0; c < K; ++c) hB[(size_t)r * K + c] = __float2bfloat16(r == c ? 1.f : 0.f);
    __nv_bfloat16 *dA, *dB; float* dOut;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dOut, 128 * 64 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap mA, mB;
    const CUtensorMapSwizzle sw = rowbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t dimsA[2] = {(cuuint64_t)K, (cuuint64_t)rowsA}, strA[1] = {(cuuint64_t)rowbytes};
    cuuint32_t boxA[2] = {(cuuint32_t)K, 256}, es[2] = {1, 1};
    // rowsA = 512 > 256 box limit: load the first 256 rows only... use two boxes via one map of 256 rows
    CUresult r1 = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, strA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dimsB[2] = {(cuuint64_t)K, (cuuint64_t)K}; cuuint32_t boxB[2] = {(cuuint32_t)K, (cuuint32_t)K};
    CUresult r2 = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, strA, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("B: rowbytes %d encode %d %d\n", rowbytes, (int)r1, (int)r2);
    CK(cudaFuncSetAttribute(probeB, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    std::vector<float> hOut(128 * 64);
    for (int shift : {0, 1, 3, 5, 8, 11}) for (int pitch : {8, 10, 12}) {
      CK(cudaMemset(dOut, 0, 128 * 64 * 4));
      probeB<<<1, 128, 80 * 1024>>>(mA, mB, 256, rowbytes, shift, pitch, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("B: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost));
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m) {
        const int row = shift + (m / 8) * pitch + (m % 8);
        if (row >= 256) continue;
        for (int c = 0; c < K; ++c) {
          const float want = __bfloat162float(hA[(size_t)row * K + c]);
          if (hOut[m * 64 + c] != want) { ++bad; if (first < 0) first = m * 64 + c; }
        }
      }
      printf("B: rowbytes %d shift %2d pitch %2d: %s (bad %d", rowbytes, shift, pitch, bad ? "MISMATCH" : "ok", bad);
      if (bad) printf(", first m=%d c=%d got %.0f", first / 64, first % 64, hOut[first]);
      printf(")\n");
    }
  }

  // ---------------- C: TMA bandwidth
  {
    int dev_clk = 0; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
    long long* dclk; CK(cudaMalloc(&dclk, 148 * 8));
    CK(cudaFuncSetAttribute(probeC, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    struct Case { const char* name; int N, H, W, C; int bw, bh, stages; };
    const Case cases[] = {
        {"L2-resident 32MB, box 18x34 (78KB) x2", 4, 128, 128, 256, 16, 32, 2},
        {"streaming 537MB, box 18x34 (78KB) x2", 16, 256, 256, 256, 16, 32, 2},
        {"L2-resident 32MB, box 10x18 (23KB) x8", 4, 128, 128, 256, 8, 16, 8},
        {"streaming 537MB, box 10x18 (23KB) x8", 16, 256, 256, 256, 8, 16, 8},
        {"streaming 537MB, box 18x18 (41KB) x4", 16, 256, 256, 256, 16, 16, 4},
    };
    for (const Case& cs : cases) {
      __nv_bfloat16* d; const size_t elems = (size_t)cs.N * cs.H * cs.W * cs.C;
      CK(cudaMalloc(&d, elems * 2)); CK(cudaMemset(d, 0, elems * 2));
      CUtensorMap m;
      cuuint64_t dims[4] = {(cuuint64_t)cs.C, (cuuint64_t)cs.W, (cuuint64_t)cs.H, (cuuint64_t)cs.N};
      cuuint64_t str[3] = {(cuuint64_t)cs.C * 2, (cuuint64_t)cs.C * 2 * cs.W, (cuuint64_t)cs.C * 2 * cs.W * cs.H};
      cuuint32_t box[4] = {64, (cuuint32_t)(cs.bw + 2), (cuuint32_t)(cs.bh + 2), 1}, es[4] = {1, 1, 1, 1};
      CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("C: encode failed %d\n", (int)r); return 1; }
      const int box_bytes = (cs.bw + 2) * (cs.bh + 2) * 128;
      const int stage_bytes = (box_bytes + 1023) / 1024 * 1024;
      const int iters = 2000;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        probeC<<<148, 64, (size_t)cs.stages * stage_bytes + 2048>>>(m, cs.stages, stage_bytes, box_bytes, iters,
                                                                   cs.W / cs.bw, cs.H / cs.bh, cs.N, cs.C / 64, cs.bw,
                                                                   cs.bh, dclk);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long hclk[148]; CK(cudaMemcpy(hclk, dclk, sizeof(hclk), cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = hclk[i] > mx ? hclk[i] : mx;
      const double bytes = (double)box_bytes * iters;
      printf("C: %-42s: %.3f ms, %.2f TB/s chip, %.1f B/clk/SM (max clk %lld)\n", cs.name, ms,
             bytes * 148 / ms / 1e9, bytes / (double)mx, mx);
      cudaFree(d);
    }
  }

  // ---------------- D: MMA issue rate from shared memory
  {
    long long* dclk; CK(cudaMalloc(&dclk, 148 * 8));
    CK(cudaFuncSetAttribute(probeD, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int n_mma = 8192;
    for (int grid : {1, 148})
      for (int N : {16, 32, 64, 128, 256})
        for (int pitch : {8, 10, 18}) {
          const int n_acc = 512 / N > 4 ? 4 : 512 / N;
          for (int rep = 0; rep < 2; ++rep) {
            probeD<<<grid, 128, 200 * 1024>>>(N, n_mma, pitch, 9, n_acc, dclk);
            CK(cudaDeviceSynchronize());
          }
          long long hclk[148]; CK(cudaMemcpy(hclk, dclk, grid * 8, cudaMemcpyDeviceToHost));
          long long mx = 0; for (int i = 0; i < grid; ++i) mx = hclk[i] > mx ? hclk[i] : mx;
          printf("D: grid %3d N %3d pitch %2d: %.1f clk/MMA (floor %d), %.0f%% of tensor peak\n", grid, N, pitch,
                 (double)mx / n_mma, N / 2, 100.0 * (N / 2) / ((double)mx / n_mma));
        }
  }
  return 0;
}
