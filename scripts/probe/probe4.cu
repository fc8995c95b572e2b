// Probe 4: TMA box-shape cost model.  All SMs issue the same box shape in a ring; report clk per box.
#include <cstdio>
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
__device__ __forceinline__ void tma_load_5d(void* dst, const void* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// coordinate pattern: c[d] = base[d] + (t % mod[d]) * mul[d]
struct Pat { int rank; int base[5], mod[5], mul[5]; };
__global__ void __launch_bounds__(64, 1)
probe(const __grid_constant__ CUtensorMap map, const __grid_constant__ Pat pat, int stages, int stage_bytes, int box_bytes, int iters, long long* clk_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + 8;
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } fence_barrier_init(); }
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0; int t = blockIdx.x * 37;
    for (int i = 0; i < iters; ++i, t += 148) {
      mbar_wait(&empty[stage], phase ^ 1);
      int c[5]; int q = t;
      for (int d = 0; d < 5; ++d) { const int m = pat.mod[d]; c[d] = pat.base[d] + (q & (m - 1)) * pat.mul[d]; q >>= (31 - __clz(m)); }
      mbar_arrive_expect_tx(&full[stage], box_bytes);
      void* dst = smem + (size_t)stage * stage_bytes;
      if (pat.rank == 2) tma_load_2d(dst, &map, &full[stage], c[0], c[1]);
      else if (pat.rank == 4) tma_load_4d(dst, &map, &full[stage], c[0], c[1], c[2], c[3]);
      else tma_load_5d(dst, &map, &full[stage], c[0], c[1], c[2], c[3], c[4]);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) { mbar_wait(&full[stage], phase); mbar_arrive(&empty[stage]); if (++stage == stages) { stage = 0; phase ^= 1; } }
    clk_out[blockIdx.x] = clock64() - t0;
  }
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
static long long* dclk;
static int run(const char* name, int rank, void* ptr, std::vector<cuuint64_t> dims, std::vector<cuuint64_t> strides, std::vector<cuuint32_t> box, Pat pat, int stages, CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m; cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, ptr, dims.data(), strides.data(), box.data(), es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return 0; }
  size_t box_bytes = 2; for (auto b : box) box_bytes *= b;
  const int stage_bytes = (int)((box_bytes + 1023) / 1024 * 1024);
  const int iters = 1000;
  pat.rank = rank;
  for (int rep = 0; rep < 2; ++rep) { probe<<<148, 64, (size_t)stages * stage_bytes + 2048>>>(m, pat, stages, stage_bytes, (int)box_bytes, iters, dclk); CK(cudaDeviceSynchronize()); }
  long long h[148]; CK(cudaMemcpy(h, dclk, sizeof(h), cudaMemcpyDeviceToHost));
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-58s box %6zu B: %7.0f clk/box, %6.1f B/clk/SM\n", name, box_bytes, (double)mx / iters, (double)box_bytes * iters / mx);
  return 0;
}
int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q; cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q); enc = (EncodeTiledFn)fp;
  CK(cudaMalloc(&dclk, 148 * 8));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  // activation tensor N=8, H=W=128, C=256 (67 MB, L2-resident-ish) and a low-res one N=8, 64x64x256
  const int N = 8, H = 128, W = 128, C = 256;
  __nv_bfloat16* act; CK(cudaMalloc(&act, (size_t)N * H * W * C * 2)); CK(cudaMemset(act, 0, (size_t)N * H * W * C * 2));
  std::vector<cuuint64_t> d4 = {C, W, H, N}, s4 = {C * 2ull, C * 2ull * W, C * 2ull * W * H};
  auto P4 = [&](int bw, int bh) { Pat p{}; p.base[0] = 0; p.mod[0] = 4; p.mul[0] = 64; p.base[1] = -1; p.mod[1] = W / bw; p.mul[1] = bw; p.base[2] = -1; p.mod[2] = H / bh; p.mul[2] = bh; p.base[3] = 0; p.mod[3] = N; p.mul[3] = 1; p.mod[4] = 1; return p; };
  run("4D halo 10w x 18h (T=1)", 4, act, d4, s4, {64, 10, 18, 1}, P4(8, 16), 4);
  run("4D halo 18w x 18h (T=2)", 4, act, d4, s4, {64, 18, 18, 1}, P4(16, 16), 4);
  run("4D halo 34w x 18h (T=4)", 4, act, d4, s4, {64, 34, 18, 1}, P4(32, 16), 2);
  run("4D halo 66w x 18h (T=8)", 4, act, d4, s4, {64, 66, 18, 1}, P4(64, 16), 1);
  run("4D halo 34w x 18h (T=4) 1 stage", 4, act, d4, s4, {64, 34, 18, 1}, P4(32, 16), 1);
  run("4D halo 34w x 9h x2 loads-equivalent", 4, act, d4, s4, {64, 34, 9, 1}, P4(32, 8), 4);
  // C = 64 tensor (contiguous pixels): N=8,H=W=256
  {
    const int C2 = 64, H2 = 256, W2 = 256;
    std::vector<cuuint64_t> dd = {C2, W2, H2, N}, ss = {C2 * 2ull, C2 * 2ull * W2, C2 * 2ull * W2 * H2};
    Pat p{}; p.mod[0] = 1; p.base[1] = -1; p.mod[1] = W2 / 32; p.mul[1] = 32; p.base[2] = -1; p.mod[2] = H2 / 16; p.mul[2] = 16; p.mod[3] = N; p.mul[3] = 1; p.mod[4] = 1;
    run("4D halo 34w x 18h, C=64 tensor (contiguous rows)", 4, act, dd, ss, {64, 34, 18, 1}, p, 2);
  }
  // nearest-x2 replicating box over a low-res tensor N=8, Hl=Wl=64, C=256: dims (C, 2, Wl, 2, N*Hl)
  {
    const int Hl = 64, Wl = 64;
    std::vector<cuuint64_t> dd = {C, 2, Wl, 2, (cuuint64_t)N * Hl}, ss = {0, C * 2ull, 0, C * 2ull * Wl};
    Pat p{}; p.mod[0] = 4; p.mul[0] = 64; p.mod[1] = 1; p.base[2] = -1; p.mod[2] = Wl / 16; p.mul[2] = 16; p.mod[3] = 1; p.mod[4] = N * Hl / 8; p.mul[4] = 8;
    run("5D dup box (64,2,18,2,8) = 36w x 16h upsampled", 5, act, dd, ss, {64, 2, 18, 2, 8}, p, 2);
    run("5D dup box (64,2,10,2,8) = 20w x 16h upsampled", 5, act, dd, ss, {64, 2, 10, 2, 8}, p, 4);
    // x-dup only, 4D: dims (C, 2, Wl, N*Hl)
    std::vector<cuuint64_t> d2 = {C, 2, Wl, (cuuint64_t)N * Hl}, s2 = {0, C * 2ull, C * 2ull * Wl};
    Pat p2{}; p2.mod[0] = 4; p2.mul[0] = 64; p2.mod[1] = 1; p2.base[2] = -1; p2.mod[2] = Wl / 16; p2.mul[2] = 16; p2.mod[3] = 32; p2.mul[3] = 9; p2.mod[4] = 1;
    run("4D x-dup box (64,2,18,9) = 36w x 9h", 4, act, d2, s2, {64, 2, 18, 9}, p2, 4);
  }
  // weights: 2D [rows][64] contiguous, box (64, bn)
  for (int bn : {64, 128, 256}) {
    std::vector<cuuint64_t> dd = {64, 65536}, ss = {128};
    Pat p{}; p.mod[0] = 1; p.mod[1] = 65536 / bn; p.mul[1] = bn; p.mod[2] = p.mod[3] = p.mod[4] = 1;
    char nm[64]; snprintf(nm, sizeof nm, "2D weights contiguous box (64, %d)", bn);
    run(nm, 2, act, dd, ss, {64, (cuuint32_t)bn}, p, 8);
  }
  // weights with row stride 9*256*2 B (gen-1 layout): box (64, 64)
  {
    std::vector<cuuint64_t> dd = {2304, 4096}, ss = {4608};
    Pat p{}; p.mod[0] = 36; p.mul[0] = 64; p.mod[1] = 64; p.mul[1] = 64; p.mod[2] = p.mod[3] = p.mod[4] = 1;
    run("2D weights strided rows box (64, 64)", 2, act, dd, ss, {64, 64}, p, 8);
  }
  // thin layers: C=16 tensor 512x512, 32-byte rows; C=32, 64-byte rows
  {
    const int C2 = 16, H2 = 512, W2 = 512;
    std::vector<cuuint64_t> dd = {C2, W2, H2, N}, ss = {C2 * 2ull, C2 * 2ull * W2, C2 * 2ull * W2 * H2};
    Pat p{}; p.mod[0] = 1; p.base[1] = -1; p.mod[1] = W2 / 32; p.mul[1] = 32; p.base[2] = -1; p.mod[2] = H2 / 16; p.mul[2] = 16; p.mod[3] = N; p.mul[3] = 1; p.mod[4] = 1;
    run("4D halo 34w x 18h, C=16 (32B rows)", 4, act, dd, ss, {16, 34, 18, 1}, p, 8, CU_TENSOR_MAP_SWIZZLE_32B);
    run("4D halo 66w x 18h, C=16 (32B rows)", 4, act, dd, ss, {16, 66, 18, 1}, p, 8, CU_TENSOR_MAP_SWIZZLE_32B);
    std::vector<cuuint64_t> d3 = {32, 512, 256, N}, s3 = {64, 64ull * 512, 64ull * 512 * 256};
    run("4D halo 34w x 18h, C=32 (64B rows)", 4, act, d3, s3, {32, 34, 18, 1}, p, 8, CU_TENSOR_MAP_SWIZZLE_64B);
  }
  return 0;
}
