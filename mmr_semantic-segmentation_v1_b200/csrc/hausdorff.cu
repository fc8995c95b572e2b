// Per-image, per-class Hausdorff distance between the predicted and the labelled mask on the device.
//
// Replaces the reference's every-25-epochs loop (SU/ModelTraining.py:625-649, 765-789): per image, one_hot ->
// .cpu() -> per class skimage.metrics.hausdorff_distance(seg_slice, label_slice) (two cKDTree nearest-neighbour
// queries between the nonzero pixel sets A = {pred == c} and B = {label == c}):
//     H(A, B) = max( max_{b in B} min_{a in A} |a - b|,  max_{a in A} min_{b in B} |a - b| )
// with H = 0 when both sets are empty and H = inf when exactly one is.
//
// Exact integer arithmetic: the kernels produce the maximal SQUARED distance (an integer), the host takes the
// square root in float64 -- the same value the cKDTree path returns.  Separable Euclidean distance transform:
//   1. column pass: g[set][c][y][x] = vertical distance from (y, x) to the nearest pixel of class c of `set`
//      in column x (65535 = none), one thread per (set, class, column);
//   2. row pass, only at the pixels of the OTHER set's class-c mask: d2 = min_x' (x - x')^2 + g[y][x']^2, searched
//      outward from x' = x and stopped as soon as (x - x')^2 >= d2; since every pixel belongs to exactly one
//      class of each map, the work is H * W searches per direction and image, independent of the class count.
// HBM traffic: the two maps once per pass plus the 2 * C * H * W uint16 table written and read once.
#include "common.h"

namespace mmr {

constexpr int kHdThreads = 256;
constexpr unsigned int kHdNone = 65535u;

template <typename TP>
__device__ __forceinline__ long long hd_value(const void* p, size_t i) {
  return (long long)reinterpret_cast<const TP*>(p)[i];
}

// grid: (ceil(W / threads), C, 2 * images of the chunk)
__global__ void __launch_bounds__(kHdThreads)
hausdorff_columns_kernel(const void* __restrict__ pred, int pred_u8, const long long* __restrict__ labels, int n0,
                         int C, int H, int W, unsigned short* __restrict__ g) {
  pdl_prologue();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int c = blockIdx.y, set = blockIdx.z & 1, img = blockIdx.z >> 1;
  const size_t base = (size_t)(n0 + img) * H * W;
  unsigned short* out = g + (((size_t)img * 2 + set) * C + c) * (size_t)H * W + x;
  unsigned int run = kHdNone;
  for (int y = 0; y < H; ++y) {   // distance to the nearest hit above
    const size_t i = base + (size_t)y * W + x;
    const long long v = set ? labels[i] : (pred_u8 ? hd_value<unsigned char>(pred, i) : hd_value<long long>(pred, i));
    run = v == c ? 0u : (run >= kHdNone ? kHdNone : run + 1u);
    out[(size_t)y * W] = (unsigned short)run;
  }
  run = kHdNone;
  for (int y = H - 1; y >= 0; --y) {   // ... and below
    const unsigned int up = out[(size_t)y * W];
    run = up == 0u ? 0u : (run >= kHdNone ? kHdNone : run + 1u);
    if (run < up) out[(size_t)y * W] = (unsigned short)run;
  }
}

// grid: (H, 2 * images of the chunk); dynamic shared memory: C * W uint16 (row y of the searched set's table)
__global__ void __launch_bounds__(kHdThreads)
hausdorff_rows_kernel(const void* __restrict__ pred, int pred_u8, const long long* __restrict__ labels, int n0, int C,
                      int H, int W, const unsigned short* __restrict__ g, unsigned long long* __restrict__ hd2) {
  pdl_prologue();
  extern __shared__ unsigned short row[];   // [C][W]
  __shared__ unsigned long long best_c[16];
  const int y = blockIdx.x, dir = blockIdx.y & 1, img = blockIdx.y >> 1;
  // dir 0: points of the label mask searched in the prediction's table (set 0); dir 1: the reverse
  const int searched = dir;          // set index of the table we search: dir 0 -> pred (0), dir 1 -> label (1)
  const unsigned short* tab = g + ((size_t)img * 2 + searched) * C * (size_t)H * W + (size_t)y * W;
  for (int k = threadIdx.x; k < C * W; k += blockDim.x) row[k] = tab[(size_t)(k / W) * H * W + (k % W)];
  if (threadIdx.x < 16) best_c[threadIdx.x] = 0ull;
  __syncthreads();
  const size_t base = (size_t)(n0 + img) * H * W + (size_t)y * W;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const long long v = dir ? (pred_u8 ? hd_value<unsigned char>(pred, base + x) : hd_value<long long>(pred, base + x))
                            : labels[base + x];
    if (v < 0 || v >= C) continue;
    const unsigned short* r = row + (size_t)v * W;
    unsigned long long best = ~0ull;
    const int rmax = max(x, W - 1 - x);
    for (int d = 0; d <= rmax; ++d) {
      const unsigned long long dd = (unsigned long long)d * d;
      if (dd >= best) break;
      if (x - d >= 0) {
        const unsigned int gv = r[x - d];
        if (gv != kHdNone) best = min(best, dd + (unsigned long long)gv * gv);
      }
      if (d && x + d < W) {
        const unsigned int gv = r[x + d];
        if (gv != kHdNone) best = min(best, dd + (unsigned long long)gv * gv);
      }
    }
    atomicMax(&best_c[(int)v], best);
  }
  __syncthreads();
  if (threadIdx.x < C && best_c[threadIdx.x])
    atomicMax(hd2 + (size_t)(n0 + img) * C + threadIdx.x, best_c[threadIdx.x]);
}

}  // namespace mmr

using namespace mmr;

extern "C" int64_t mmr_hausdorff_workspace_bytes(int C, int H, int W) {
  return (int64_t)2 * C * H * W * (int64_t)sizeof(unsigned short);
}

extern "C" int mmr_hausdorff_sq(const void* pred, int pred_u8, const int64_t* labels, int N, int C, int H, int W,
                                void* workspace, int64_t workspace_bytes, unsigned long long* hd2,
                                mmr_stream_t stream) {
  MMR_REQUIRE(pred && labels && workspace && hd2, "null argument");
  MMR_REQUIRE(N >= 1 && C >= 1 && C <= 16 && H >= 1 && W >= 1 && H < 65535 && W < 65535,
              "hausdorff: need 1 <= classes <= 16 and H, W < 65535 (got C=%d H=%d W=%d)", C, H, W);
  const int64_t per_image = mmr_hausdorff_workspace_bytes(C, H, W);
  MMR_REQUIRE(workspace_bytes >= per_image, "hausdorff: workspace of %lld bytes, one image needs %lld",
              (long long)workspace_bytes, (long long)per_image);
  const size_t row_smem = (size_t)C * W * sizeof(unsigned short);
  MMR_REQUIRE(row_smem <= 200 * 1024, "hausdorff: classes * W * 2 bytes = %zu exceed the shared-memory row buffer", row_smem);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(hausdorff_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  MMR_CUDA_CHECK(cudaMemsetAsync(hd2, 0, sizeof(unsigned long long) * (size_t)N * C, as_stream(stream)));
  int chunk = (int)(workspace_bytes / per_image);
  if (chunk > 16384) chunk = 16384;   // gridDim.z / gridDim.y limits (2 * chunk <= 65535)
  for (int n0 = 0; n0 < N; n0 += chunk) {
    const int imgs = N - n0 < chunk ? N - n0 : chunk;
    dim3 g1((W + kHdThreads - 1) / kHdThreads, C, 2 * imgs);
    mmr_launch((hausdorff_columns_kernel), g1, dim3(kHdThreads), 0, as_stream(stream), pred, pred_u8,
               reinterpret_cast<const long long*>(labels), n0, C, H, W, reinterpret_cast<unsigned short*>(workspace));
    MMR_CUDA_CHECK(cudaGetLastError());
    dim3 g2(H, 2 * imgs);
    mmr_launch((hausdorff_rows_kernel), g2, dim3(kHdThreads), row_smem, as_stream(stream), pred, pred_u8,
               reinterpret_cast<const long long*>(labels), n0, C, H, W,
               reinterpret_cast<const unsigned short*>(workspace), hd2);
    MMR_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}
