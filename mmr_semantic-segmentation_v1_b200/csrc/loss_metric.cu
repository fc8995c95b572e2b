// Loss and metric kernels (HBM-bound, one pass over the logits each):
//   * softmax + soft-Dice + cross-entropy forward (per-(image,class) sums by warp-shuffle
//     reduction, deterministic two-level combine) and backward (dlogits in one pass);
//   * argmax -> per-image confusion matrix with shared-memory integer atomics (bit-exact).
// Logits are fp32 NCHW (what `model(img)` returns in the reference), labels int64.
//
// Reference arithmetic restated here (see oracle/ for the CPU restatement and citations):
//   SU/dice_loss.py:118-159   p = softmax(x,1); y = one_hot(t) + 1e-6 (kornia);
//                             I = sum_hw p*y; Card = sum_hw (p + y);
//                             dice = mean_{n,c} (1 - (2I + eps)/(Card + eps))
//   SU/ModelTraining.py:600-603  loss = w*dice + (1-w)*CrossEntropy(x, t)
//   monai DiceCELoss(softmax=True) (ED/Main_MMR_SegModel.py:578,709): same with y exact one-hot
//                             and smoothing 1e-5 in numerator and denominator, weights 1 and 1.
#include <atomic>

#include "common.h"
#include "ptx.cuh"

namespace mmr {

constexpr int kLossThreads = 256;
constexpr int kMaxClasses = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// workspace layout (doubles):
//   [0, N*C*3)                    reduced I, P, Y per (n,c)
//   [N*C*3, +4)                   ce_sum, ce_count, dice_loss, total_loss
//   then N*nblk*(3C+2)            per-block partials: for each (n, blk): I[C], P[C], Y[C], ce, cnt
__host__ __device__ inline int64_t ws_head_doubles(int N, int C) { return (int64_t)N * C * 3 + 4; }

constexpr int kLossTickets = 64;
__device__ unsigned int g_loss_tickets[kLossTickets];   // self-resetting arrival counters, one per launch in flight

// The combine step of the forward, run by the LAST block of dice_ce_fwd_kernel (ticket): one launch for the whole
// forward.  Partials written by the other blocks are read with ld.global.cg.
__device__ __forceinline__ void dice_ce_finalize_body(double* __restrict__ ws, int N, int C, int nblk,
                                                      const MmrLossParams& prm, float* __restrict__ out) {
  double* red = ws;
  double* tail = red + (size_t)N * C * 3;
  const double* part = ws + ws_head_doubles(N, C);
  __shared__ double sh_dice[kLossThreads];
  __shared__ double sh_ce[kLossThreads], sh_cnt[kLossThreads];
  double dice_acc = 0.0, ce_acc = 0.0, cnt_acc = 0.0;
  // one warp per (image, class): the lanes stride over the block partials (loads in flight instead of a serial
  // chain of L2 round trips), then a fixed-order shuffle tree
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int k = warp; k < N * C; k += nwarp) {
    const int n = k / C, c = k % C;
    double I = 0.0, P = 0.0, Y = 0.0;
    for (int b = lane; b < nblk; b += 32) {
      const double* src = part + ((size_t)n * nblk + b) * (3 * C + 2);
      I += __ldcg(src + c);
      P += __ldcg(src + C + c);
      Y += __ldcg(src + 2 * C + c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      I += __shfl_xor_sync(0xffffffffu, I, o);
      P += __shfl_xor_sync(0xffffffffu, P, o);
      Y += __shfl_xor_sync(0xffffffffu, Y, o);
    }
    if (lane == 0) {
      red[(size_t)k * 3 + 0] = I;
      red[(size_t)k * 3 + 1] = P;
      red[(size_t)k * 3 + 2] = Y;
      if (c < prm.dice_channels)
        dice_acc += 1.0 - (2.0 * I + (double)prm.dice_eps_nr) / (P + Y + (double)prm.dice_eps_dr);
    }
  }
  for (int k = threadIdx.x; k < N * nblk; k += blockDim.x) {
    const double* src = part + (size_t)k * (3 * C + 2);
    ce_acc += __ldcg(src + 3 * C);
    cnt_acc += __ldcg(src + 3 * C + 1);
  }
  sh_dice[threadIdx.x] = dice_acc;
  sh_ce[threadIdx.x] = ce_acc;
  sh_cnt[threadIdx.x] = cnt_acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh_dice[threadIdx.x] += sh_dice[threadIdx.x + s];
      sh_ce[threadIdx.x] += sh_ce[threadIdx.x + s];
      sh_cnt[threadIdx.x] += sh_cnt[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double dice = sh_dice[0] / ((double)N * (double)prm.dice_channels);
    const double cnt = sh_cnt[0];
    const double cem = cnt > 0.0 ? sh_ce[0] / cnt : 0.0;
    tail[0] = sh_ce[0];
    tail[1] = cnt;
    tail[2] = dice;
    tail[3] = (double)prm.w_dice * dice + (double)prm.w_ce * cem;
    out[0] = (float)tail[3];
    out[1] = (float)dice;
    out[2] = (float)cem;
  }
}

// register cap: the combine step inlined at the end would otherwise set the allocation (98-120 registers for a
// main loop that needs 40-80) and halve the warps that hide the loop's exp / divide chains
template <int CP>
__global__ void __launch_bounds__(kLossThreads, CP <= 4 ? 4 : (CP <= 10 ? 3 : 1))
dice_ce_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int C,
                   int64_t HW, MmrLossParams prm, double* __restrict__ ws, int nblk, float* __restrict__ result,
                   int ticket_slot) {
  pdl_prologue();
  const int n = blockIdx.y;
  const float* lg = logits + (size_t)n * C * HW;
  const int64_t* lb = labels + (size_t)n * HW;
  float accI[CP], accP[CP], accY[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) accI[c] = accP[c] = accY[c] = 0.f;
  float ce = 0.f, cnt = 0.f;
  // U pixels per iteration, all their loads issued before the first use (one pixel = C + 1 loads in flight per
  // thread left the 10-class launches latency-bound); the pixels are still accumulated in index order
  constexpr int U = CP <= 10 ? 2 : 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < HW; i0 += U * stride) {
    float vv[U][CP];
    int64_t tt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      const bool on = i < HW;
#pragma unroll
      for (int c = 0; c < CP; ++c) vv[u][c] = (on && c < C) ? __ldg(lg + (size_t)c * HW + i) : -INFINITY;
      tt[u] = on ? __ldg(lb + i) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * stride >= HW) break;
      float (&v)[CP] = vv[u];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < CP; ++c) mx = fmaxf(mx, v[c]);
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        v[c] = c < C ? expf(v[c] - mx) : 0.f;
        se += v[c];
      }
      const float inv = 1.f / se;
      const int64_t t = tt[u];
      const int ti = (t >= 0 && t < C) ? (int)t : -1;
      float pt = 0.f;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const float p = v[c] * inv;
        const float y = (c == ti ? 1.f : 0.f) + prm.onehot_eps;
        accI[c] += p * y;
        accP[c] += p;
        accY[c] += y;
        if (c == ti) pt = p;
      }
      if (t != prm.ce_ignore_index && ti >= 0) {
        ce += -logf(fmaxf(pt, 1e-38f));
        cnt += 1.f;
      }
    }
  }
  __shared__ float sh[kLossThreads / 32][3 * CP + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < CP; ++c) {
    const float a = warp_sum(accI[c]), b = warp_sum(accP[c]), d = warp_sum(accY[c]);
    if (lane == 0) sh[warp][c] = a, sh[warp][CP + c] = b, sh[warp][2 * CP + c] = d;
  }
  ce = warp_sum(ce);
  cnt = warp_sum(cnt);
  if (lane == 0) sh[warp][3 * CP] = ce, sh[warp][3 * CP + 1] = cnt;
  __syncthreads();
  double* out = ws + ws_head_doubles(gridDim.y, C) + ((size_t)n * nblk + blockIdx.x) * (3 * C + 2);
  for (int k = threadIdx.x; k < 3 * C + 2; k += blockDim.x) {
    // map compact index k (C-strided) to padded shared index (CP-strided)
    int sidx;
    if (k < 3 * C)
      sidx = (k / C) * CP + (k % C);
    else
      sidx = 3 * CP + (k - 3 * C);
    double s = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) s += (double)sh[w][sidx];
    out[k] = s;
  }
  // the last block to get here owns every partial: it combines them (no second launch)
  __shared__ unsigned int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
    is_last = atomicAdd(&g_loss_tickets[ticket_slot], 1u) == gridDim.x * gridDim.y - 1 ? 1u : 0u;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  dice_ce_finalize_body(ws, (int)gridDim.y, C, nblk, prm, result);
  if (threadIdx.x == 0) g_loss_tickets[ticket_slot] = 0u;
}

// dL/dp[n,c,i] = -w_dice/(N*Cd) * (2*y - D_nc)/(Card_nc + eps_dr)  with D = (2I+eps_nr)/(Card+eps_dr)
// dL/dz = p * (g - sum_k g_k p_k)  +  w_ce * (p - onehot)/count
template <int CP>
__global__ void __launch_bounds__(kLossThreads)
dice_ce_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int N, int C,
                   int64_t HW, MmrLossParams prm, const double* __restrict__ ws,
                   float grad_scale, const float* __restrict__ grad_scale_dev,
                   float* __restrict__ dlogits) {
  pdl_prologue();
  if (grad_scale_dev) grad_scale *= __ldg(grad_scale_dev);
  const int n = blockIdx.y;
  const double* red = ws;
  const double* tail = red + (size_t)N * C * 3;
  __shared__ float coefA[CP], coefD[CP];  // a = w_dice/(N*Cd)/(Card+eps), D
  if (threadIdx.x < CP) {
    const int c = threadIdx.x;
    float a = 0.f, D = 0.f;
    if (c < C && c < prm.dice_channels) {
      const double I = red[((size_t)n * C + c) * 3 + 0];
      const double card = red[((size_t)n * C + c) * 3 + 1] + red[((size_t)n * C + c) * 3 + 2];
      const double den = card + (double)prm.dice_eps_dr;
      D = (float)((2.0 * I + (double)prm.dice_eps_nr) / den);
      a = (float)((double)prm.w_dice / ((double)N * (double)prm.dice_channels) / den);
    }
    coefA[c] = a;
    coefD[c] = D;
  }
  __syncthreads();
  const double cnt = tail[1];
  const float ce_w = cnt > 0.0 ? (float)((double)prm.w_ce / cnt) : 0.f;
  const float* lg = logits + (size_t)n * C * HW;
  const int64_t* lb = labels + (size_t)n * HW;
  float* dl = dlogits + (size_t)n * C * HW;
  constexpr int U = CP <= 10 ? 2 : 1;      // as in the forward: U pixels' loads in flight per thread
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < HW; i0 += U * stride) {
    float pp[U][CP];
    int64_t tt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      const bool on = i < HW;
#pragma unroll
      for (int c = 0; c < CP; ++c) pp[u][c] = (on && c < C) ? __ldg(lg + (size_t)c * HW + i) : -INFINITY;
      tt[u] = on ? __ldg(lb + i) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= HW) break;
      float (&p)[CP] = pp[u];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < CP; ++c) mx = fmaxf(mx, p[c]);
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        p[c] = c < C ? expf(p[c] - mx) : 0.f;
        se += p[c];
      }
      const float inv = 1.f / se;
      const int64_t t = tt[u];
      const int ti = (t >= 0 && t < C) ? (int)t : -1;
      float g[CP];
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        p[c] *= inv;
        const float y = (c == ti ? 1.f : 0.f) + prm.onehot_eps;
        g[c] = -coefA[c] * (2.f * y - coefD[c]);
        dot += g[c] * p[c];
      }
      const bool ce_on = t != prm.ce_ignore_index && ti >= 0;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        if (c < C) {
          float d = p[c] * (g[c] - dot);
          if (ce_on) d += ce_w * (p[c] - (c == ti ? 1.f : 0.f));
          dl[(size_t)c * HW + i] = d * grad_scale;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ confusion matrix
template <int CP>
__global__ void __launch_bounds__(kLossThreads)
confusion_logits_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int C,
                        int64_t HW, unsigned long long* __restrict__ cm,
                        int64_t* __restrict__ pred_out) {
  pdl_prologue();
  __shared__ unsigned int hist[kMaxClasses * kMaxClasses];
  for (int k = threadIdx.x; k < C * C; k += blockDim.x) hist[k] = 0u;
  __syncthreads();
  const int n = blockIdx.y;
  const float* lg = logits + (size_t)n * C * HW;
  const int64_t* lb = labels ? labels + (size_t)n * HW : nullptr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW;
       i += (int64_t)gridDim.x * blockDim.x) {
    float best = __ldg(lg + i);
    int arg = 0;
#pragma unroll
    for (int c = 1; c < CP; ++c) {
      if (c < C) {
        const float v = __ldg(lg + (size_t)c * HW + i);
        // torch.argmax: first maximal index; NaN counts as maximal
        if (v > best || (v != v && best == best)) best = v, arg = c;
      }
    }
    if (pred_out) pred_out[(size_t)n * HW + i] = arg;
    if (lb) {
      const int64_t t = __ldg(lb + i);
      if (t >= 0 && t < C) atomicAdd(&hist[(int)t * C + arg], 1u);
    }
  }
  __syncthreads();
  if (cm) {
    for (int k = threadIdx.x; k < C * C; k += blockDim.x)
      if (hist[k]) atomicAdd(&cm[(size_t)n * C * C + k], (unsigned long long)hist[k]);
  }
}

// labels[n][i] = first channel holding the maximum of a one-hot map (torch.argmax rule).
template <typename T>
__global__ void __launch_bounds__(kLossThreads)
onehot_to_labels_kernel(const T* __restrict__ oh, int C, int64_t HW, int64_t* __restrict__ labels) {
  pdl_prologue();
  const int n = blockIdx.y;
  const T* src = oh + (size_t)n * C * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW;
       i += (int64_t)gridDim.x * blockDim.x) {
    T best = src[i];
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      const T v = src[(size_t)c * HW + i];
      if (v > best) best = v, arg = c;
    }
    labels[(size_t)n * HW + i] = arg;
  }
}

// overflow != 0: one extra bin (index C) collects out-of-range predictions / labels, so that
// row and column sums are the plain histograms smp's get_stats uses; cm is then (C+1) x (C+1).
__global__ void __launch_bounds__(kLossThreads)
confusion_preds_kernel(const int64_t* __restrict__ preds, const int64_t* __restrict__ labels, int C,
                       int64_t HW, int64_t ignore_index, int overflow,
                       unsigned long long* __restrict__ cm) {
  pdl_prologue();
  __shared__ unsigned int hist[(kMaxClasses + 1) * (kMaxClasses + 1)];
  const int Cb = C + (overflow ? 1 : 0);
  for (int k = threadIdx.x; k < Cb * Cb; k += blockDim.x) hist[k] = 0u;
  __syncthreads();
  const int n = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = __ldg(preds + (size_t)n * HW + i), t = __ldg(labels + (size_t)n * HW + i);
    if (t == ignore_index) continue;
    const bool t_ok = t >= 0 && t < C, p_ok = p >= 0 && p < C;
    if (overflow) {
      if (!t_ok) t = C;
      if (!p_ok) p = C;
    } else if (!t_ok || !p_ok) {
      continue;
    }
    atomicAdd(&hist[(int)t * Cb + (int)p], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < Cb * Cb; k += blockDim.x)
    if (hist[k]) atomicAdd(&cm[(size_t)n * Cb * Cb + k], (unsigned long long)hist[k]);
}

static int blocks_per_image(int64_t HW, int N) {
  // keep each block under 2^32 counts and the grid around 4 CTAs per SM
  int64_t b = (HW + kLossThreads * 8 - 1) / (kLossThreads * 8);
  const int64_t cap = ((int64_t)num_sms() * 4 + N - 1) / N;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace mmr

using namespace mmr;

#define DISPATCH_CP(C, CALL)                                   \
  do {                                                         \
    if ((C) <= 2) { constexpr int CP = 2; CALL; }              \
    else if ((C) <= 4) { constexpr int CP = 4; CALL; }         \
    else if ((C) <= 8) { constexpr int CP = 8; CALL; }         \
    else if ((C) <= 10) { constexpr int CP = 10; CALL; }       \
    else if ((C) <= 16) { constexpr int CP = 16; CALL; }       \
    else { constexpr int CP = 32; CALL; }                      \
  } while (0)

extern "C" int64_t mmr_dice_ce_workspace_doubles(int N, int C, int nblk) {
  return ws_head_doubles(N, C) + (int64_t)N * nblk * (3 * C + 2);
}

extern "C" int mmr_dice_ce_fwd(const float* logits, const int64_t* labels, int N, int C, int H, int W,
                               const MmrLossParams* p, double* workspace, int nblk, float* out,
                               mmr_stream_t stream) {
  MMR_REQUIRE(C >= 1 && C <= kMaxClasses, "classes must be in [1,%d], got %d", kMaxClasses, C);
  MMR_REQUIRE(p && p->dice_channels >= 1 && p->dice_channels <= C, "dice_channels must be in [1,C]");
  MMR_REQUIRE(N * C <= 65536, "N*C too large for the finalize kernel");
  const int64_t HW = (int64_t)H * W;
  dim3 grid(nblk, N);
  // one self-resetting arrival counter per launch in flight: a launch takes the next slot of a ring of 64
  static std::atomic<unsigned> next_ticket{0};
  const int slot = (int)(next_ticket.fetch_add(1u) % kLossTickets);
  DISPATCH_CP(C, (mmr_launch((dice_ce_fwd_kernel<CP>), grid, kLossThreads, 0, as_stream(stream), logits, labels, C, HW, *p, workspace, nblk, out, slot)));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_dice_ce_bwd(const float* logits, const int64_t* labels, int N, int C, int H, int W,
                               const MmrLossParams* p, const double* workspace, float grad_scale,
                               const float* grad_scale_dev, float* dlogits, mmr_stream_t stream) {
  MMR_REQUIRE(C >= 1 && C <= kMaxClasses, "classes must be in [1,%d], got %d", kMaxClasses, C);
  const int64_t HW = (int64_t)H * W;
  dim3 grid(blocks_per_image(HW, N), N);
  DISPATCH_CP(C, (mmr_launch((dice_ce_bwd_kernel<CP>), grid, kLossThreads, 0, as_stream(stream), logits, labels, N, C, HW, *p, workspace, grad_scale, grad_scale_dev, dlogits)));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_confusion_from_logits(const float* logits, const int64_t* labels, int N, int C,
                                         int H, int W, int64_t* cm, int64_t* pred_out,
                                         mmr_stream_t stream) {
  MMR_REQUIRE(C >= 1 && C <= kMaxClasses, "classes must be in [1,%d], got %d", kMaxClasses, C);
  MMR_REQUIRE(cm == nullptr || labels != nullptr, "confusion matrix requested without labels");
  const int64_t HW = (int64_t)H * W;
  dim3 grid(blocks_per_image(HW, N), N);
  DISPATCH_CP(C, (mmr_launch((confusion_logits_kernel<CP>), grid, kLossThreads, 0, as_stream(stream), logits, labels, C, HW, reinterpret_cast<unsigned long long*>(cm), pred_out)));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_onehot_to_labels(const void* onehot, int is_float, int N, int C, int H, int W,
                                    int64_t* labels, mmr_stream_t stream) {
  const int64_t HW = (int64_t)H * W;
  dim3 grid(blocks_per_image(HW, N), N);
  if (is_float)
    mmr_launch((onehot_to_labels_kernel<float>), grid, kLossThreads, 0, as_stream(stream), reinterpret_cast<const float*>(onehot), C, HW, labels);
  else
    mmr_launch((onehot_to_labels_kernel<int64_t>), grid, kLossThreads, 0, as_stream(stream), reinterpret_cast<const int64_t*>(onehot), C, HW, labels);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_confusion_from_preds(const int64_t* preds, const int64_t* labels, int N, int C,
                                        int64_t npix, int64_t ignore_index, int overflow_bin,
                                        int64_t* cm, mmr_stream_t stream) {
  MMR_REQUIRE(C >= 1 && C <= kMaxClasses, "classes must be in [1,%d], got %d", kMaxClasses, C);
  dim3 grid(blocks_per_image(npix, N), N);
  mmr_launch((confusion_preds_kernel), grid, kLossThreads, 0, as_stream(stream), preds, labels, C, npix, ignore_index, overflow_bin, reinterpret_cast<unsigned long long*>(cm));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}
