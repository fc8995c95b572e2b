// Probe: can a K-major SWIZZLE_128B UMMA A-operand be a *shifted, strided view* of a TMA-written
// halo tile?  rows of 128 B (64 bf16), 8-row groups strided by SBO = pitch*128 B, start address
// = base + shift*128 B (not 1024-aligned).  D = A_view * I, compared on the host.
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

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int rowsA,
      int shift, int pitch, int base_off_mode, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                    // rowsA * 128 B
  uint8_t* sB = smem + 48 * 1024;        // 64 x 128 B
  uint64_t* bar = (uint64_t*)(smem + 64 * 1024);
  uint64_t* bar2 = bar + 1;
  uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tptr, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, rowsA * 128 + 64 * 128);
    tma_load_2d(sA, &mapA, bar, 0, 0);
    tma_load_2d(sB, &mapB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    const uint32_t a0 = smem_u32(sA) + shift * 128;
    const uint32_t b0 = smem_u32(sB);
    for (int k = 0; k < 4; ++k) {
      uint64_t da = make_smem_desc(a0 + k * 32, 16, pitch * 128, 2);
      if (base_off_mode == 1) da |= (uint64_t)((a0 >> 7) & 7) << 49;
      const uint64_t db = make_smem_desc(b0 + k * 32, 16, 1024, 2);
      umma_bf16(tb, da, db, idesc, k != 0);
    }
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tb, 64);
}

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int rowsA = 256;
  std::vector<__nv_bfloat16> hA(rowsA * 64), hB(64 * 64);
  for (int r = 0; r < rowsA; ++r) for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2bfloat16((float)(r + c * 0.001f * 0 + (c % 7) * 256));
  for (int r = 0; r < 64; ++r) for (int c = 0; c < 64; ++c) hB[r * 64 + c] = __float2bfloat16(r == c ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB; float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mA, mB;
  cuuint64_t dimsA[2] = {64, (cuuint64_t)rowsA}, strA[1] = {128}; cuuint32_t boxA[2] = {64, (cuuint32_t)rowsA}, es[2] = {1, 1};
  CUresult r1 = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, strA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cuuint64_t dimsB[2] = {64, 64}; cuuint32_t boxB[2] = {64, 64};
  CUresult r2 = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, strA, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d %d\n", (int)r1, (int)r2);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  std::vector<float> hOut(128 * 64);
  const int shifts[] = {0, 1, 3, 8, 11}; const int pitches[] = {8, 10, 9, 16, 18};
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : shifts) for (int pitch : pitches) {
      cudaMemset(dOut, 0, 128 * 64 * 4);
      probe<<<1, 128, 80 * 1024>>>(mA, mB, rowsA, shift, pitch, mode, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d pitch %d: CUDA error %s\n", mode, shift, pitch, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m) {
        const int row = shift + (m / 8) * pitch + (m % 8);
        for (int c = 0; c < 64; ++c) {
          const float want = __bfloat162float(hA[row * 64 + c]);
          if (hOut[m * 64 + c] != want) { ++bad; if (first < 0) first = m * 64 + c; }
        }
      }
      printf("mode %d shift %2d pitch %2d: %s (bad %d", mode, shift, pitch, bad ? "MISMATCH" : "ok", bad);
      if (bad) printf(", first m=%d c=%d got %.0f", first / 64, first % 64, hOut[first]);
      printf(")\n");
    }
  return 0;
}
