// 1x1 classification head over a 64-channel full-resolution feature map: the reference's
// `ResNetUNet.conv_last = nn.Conv2d(64, n_class, 1)` (SU/UArchModel/resnet_unet.py:204, applied :298) behind
// `conv_original_size2` (conv3x3 + bias + ReLU, :199-201).
//
// 64 -> n_class (<= 16) at 512 x 512 is 640-1024 MACs per pixel against 128 B of activation: HBM-bound
// (AI ~ 8 F/B), so it runs on the CUDA cores in fp32 straight from the fp32 master weights and the fp32 NCHW
// logits / logit gradients; nothing is padded to a tensor-core tile, nothing is repacked:
//
//   forward   logits[n,k,y,x] = b[k] + sum_c x[n,y,x,c] W[k,c]           x read once, logits written once
//   backward  ONE pass over x and dlogits produces
//               dx[n,y,x,c] = (x > 0) * sum_k dl[n,k,y,x] W[k,c]          (the producer's ReLU mask applied: this IS
//                                                                         the producer conv's dz; bf16 NHWC)
//               dW[k,c]     = sum_p dl[k,p] x[p,c],   db[k] = sum_p dl[k,p]
//               dbL[c]      = sum_p dx[p,c]                              (bias gradient of the producer conv)
//             replacing head_grad_prep + a tensor-core dgrad + wgrad (each padded 10 -> 16 channels) + the producer's
//             grad_gather / bias_grad_finalize: 5.4 GB of traffic -> 2.5 GB at 32 x 512 x 512.
//
// Backward kernel: persistent CTAs (two per SM) walk 256-pixel tiles; a tile's activation rows and dl values arrive
// by cp.async into double-buffered shared memory while the previous tile is processed (16-byte chunks
// XOR-swizzled by the pixel index, so that rows accessed by a quarter warp and chunk columns read in phase B are
// both conflict-free).  Phase A, one thread per pixel: dx is computed, masked and left in shared memory, from
// where whole rows are copied out per warp instruction.  Phase B, the same threads regrouped as
// (k half, 8-channel chunk, pixel slice): each accumulates a CO/2 x 8 block of dW over its sixteenth of the
// tile in registers for the whole launch, plus db and dbL.  At the end a CTA folds its sixteen slices in
// shared memory and writes ONE partial vector; the finalize kernel adds the partials in a fixed order
// (deterministic, no atomics).
#include "common.h"
#include "ptx.cuh"

namespace mmr {

constexpr int kPwThreads = 256;
constexpr int kPwTile = 256;   // pixels per tile = threads per CTA
constexpr int kPwCin = 64;

// The head's weights, zero-padded to CO rows, live in constant memory for the duration of a launch: every FFMA of the
// per-pixel products takes its weight as a constant-bank operand (from shared memory the broadcast LDS.128 reads
// bound both kernels: 160 of them per pixel at 4 clk each on the SM's one shared-memory pipe).  A one-block
// staging kernel in front of each launch copies them from the fp32 master (stream order keeps launches apart).
__constant__ float c_pw_w[16 * kPwCin];
__constant__ float c_pw_b[16];

__global__ void pointwise_head_stage_kernel(const float* __restrict__ w, const float* __restrict__ bias, int cout,
                                            float* __restrict__ w_dst, float* __restrict__ b_dst) {
  pdl_prologue();
  for (int i = threadIdx.x; i < 16 * kPwCin; i += blockDim.x) w_dst[i] = (i / kPwCin) < cout ? __ldg(w + i) : 0.f;
  if (threadIdx.x < 16) b_dst[threadIdx.x] = (threadIdx.x < cout && bias) ? __ldg(bias + threadIdx.x) : 0.f;
}

__device__ __forceinline__ void pw_unpack8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(w[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}

template <int CO>
__global__ void __launch_bounds__(kPwThreads)
pointwise_head_fwd_kernel(const __nv_bfloat16* __restrict__ x, int cout, int64_t npix, int64_t hw,
                          float* __restrict__ out) {
  pdl_prologue();
  for (int64_t p = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; p < npix; p += (int64_t)gridDim.x * kPwThreads) {
    const uint4* row = reinterpret_cast<const uint4*>(x + p * kPwCin);
    uint4 raw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) raw[j] = __ldg(row + j);
    float acc[CO];
#pragma unroll
    for (int k = 0; k < CO; ++k) acc[k] = c_pw_b[k];
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float v[8];
      pw_unpack8(raw[c8], v);
#pragma unroll
      for (int k = 0; k < CO; ++k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k] += v[j] * c_pw_w[k * kPwCin + c8 * 8 + j];
      }
    }
    const int64_t n = npix <= 0x7fffffff ? (int64_t)((uint32_t)p / (uint32_t)hw) : p / hw;
    const int64_t q = p - n * hw;
    float* o = out + n * cout * hw + q;
#pragma unroll
    for (int k = 0; k < CO; ++k)
      if (k < cout) o[(int64_t)k * hw] = acc[k];
  }
}

// partial vector of one CTA: [CO * 64] dW, [CO] db, [64] dbL
template <int CO>
__host__ __device__ constexpr int pw_partial_floats() { return CO * kPwCin + CO + kPwCin; }

template <int CO>
struct PwSmem {
  uint4 xs[2][kPwTile][8];    // activation rows (double-buffered), chunk j of pixel p at [p][j ^ (p & 7)]
  uint4 ds[kPwTile][8];       // masked dx rows, same layout
  float dls[CO][kPwTile];     // dl[k][p]
};

__device__ __forceinline__ void pw_cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;      // 0 source bytes: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}

template <int CO>
__global__ void __launch_bounds__(kPwThreads, 2)
pointwise_head_bwd_kernel(const float* __restrict__ dl, const __nv_bfloat16* __restrict__ x, int cout,
                          int relu_mask, int64_t npix, int64_t hw, __nv_bfloat16* __restrict__ dx,
                          float* __restrict__ partial) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t pw_smem_raw[];
  PwSmem<CO>& sm = *reinterpret_cast<PwSmem<CO>*>(pw_smem_raw);
  constexpr int KH = CO / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // phase B role: k half, channel chunk, pixel slice (16 slices: pixel p of a tile belongs to slice p % 16)
  const int cc = lane & 7, kh = warp & 1, ps = (warp >> 1) * 4 + (lane >> 3);
  // tile movers: thread t carries 16-byte chunk t % 8 of pixels t / 8 + 32 i: a warp instruction covers four whole
  // 128-byte rows (one thread per pixel would touch 32 lines per instruction: 8x the L1 tag cycles)
  const int mv_chunk = tid & 7, mv_row = tid >> 3;
  float accW[KH][8], accb[KH], accL[8];
#pragma unroll
  for (int k = 0; k < KH; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) accW[k][j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) accL[j] = 0.f;
  const int64_t ntiles = (npix + kPwTile - 1) / kPwTile;

  // a tile's x rows come by cp.async into the other buffer; its dl values (this thread's pixel) are prefetched
  // into registers one tile ahead (two CTAs of 2 x 32 KB + 32 KB + dl tile just fit an SM)
  auto issue_tile = [&](int64_t tile, int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 32 + mv_row;
      const int64_t p = tile * kPwTile + r;
      const bool valid = p < npix;
      pw_cp_async16(&sm.xs[buf][r][mv_chunk ^ (r & 7)], x + (valid ? p : 0) * kPwCin + mv_chunk * 8, valid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float dln[CO];
  auto fetch_dl = [&](int64_t tile) {
    const int64_t p = tile * kPwTile + tid;
    const bool valid = p < npix;
    const int64_t pp = valid ? p : 0;
    // 32-bit division where the pixel index allows it (a 64-bit one costs ~100 instructions per tile)
    const int64_t n = npix <= 0x7fffffff ? (int64_t)((uint32_t)pp / (uint32_t)hw) : pp / hw;
    const int64_t q = pp - n * hw;
    const float* d = dl + n * cout * hw + q;
#pragma unroll
    for (int k = 0; k < CO; ++k) dln[k] = (valid && k < cout) ? __ldg(d + (int64_t)k * hw) : 0.f;
  };

  int buf = 0;
  if ((int64_t)blockIdx.x < ntiles) {
    issue_tile(blockIdx.x, 0);
    fetch_dl(blockIdx.x);
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();      // this tile has landed; the previous tile's phase B and copy-out are done
    float dlv[CO];
#pragma unroll
    for (int k = 0; k < CO; ++k) {
      dlv[k] = dln[k];
      sm.dls[k][tid] = dln[k];
    }
    if (tile + gridDim.x < ntiles) {
      issue_tile(tile + gridDim.x, buf ^ 1);
      fetch_dl(tile + gridDim.x);
    }
    // ---- phase A: one pixel per thread
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float xv[8], a[8];
      pw_unpack8(sm.xs[buf][tid][c8 ^ (tid & 7)], xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
      for (int k = 0; k < CO; ++k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += dlv[k] * c_pw_w[k * kPwCin + c8 * 8 + j];
      }
      if (relu_mask) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = xv[j] > 0.f ? a[j] : 0.f;
      }
      uint4 o;
      o.x = pack_bf16x2(a[0], a[1]);
      o.y = pack_bf16x2(a[2], a[3]);
      o.z = pack_bf16x2(a[4], a[5]);
      o.w = pack_bf16x2(a[6], a[7]);
      sm.ds[tid][c8 ^ (tid & 7)] = o;
    }
    __syncthreads();
    // ---- copy the dx tile out, whole rows per warp instruction
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 32 + mv_row;
      const int64_t p = tile * kPwTile + r;
      if (p < npix) *reinterpret_cast<uint4*>(dx + p * kPwCin + mv_chunk * 8) = sm.ds[r][mv_chunk ^ (r & 7)];
    }
    // ---- phase B: dW / db / dbL over this thread's sixteenth of the tile
#pragma unroll 4
    for (int it = 0; it < kPwTile / 16; ++it) {
      const int q = it * 16 + ps;
      float xv[8];
      pw_unpack8(sm.xs[buf][q][cc ^ (q & 7)], xv);
#pragma unroll
      for (int k = 0; k < KH; ++k) {
        const float d = sm.dls[kh * KH + k][q];
        accb[k] += d;
#pragma unroll
        for (int j = 0; j < 8; ++j) accW[k][j] += d * xv[j];
      }
      if ((it & 1) == kh) {      // warp-uniform: the two k halves share the column sums of dx
        float dv[8];
        pw_unpack8(sm.ds[q][cc ^ (q & 7)], dv);
#pragma unroll
        for (int j = 0; j < 8; ++j) accL[j] += dv[j];
      }
    }
  }
  // ---- fold the sixteen pixel slices (and, for dbL, the two k halves) of this CTA, write one partial vector
  __syncthreads();
  float* red = reinterpret_cast<float*>(pw_smem_raw);      // [16][CO * 64] dW, then [16][CO] db, then [32][64] dbL
  float* red_b = red + 16 * CO * kPwCin;
  float* red_l = red_b + 16 * CO;
#pragma unroll
  for (int k = 0; k < KH; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[(ps * CO + kh * KH + k) * kPwCin + cc * 8 + j] = accW[k][j];
    if (cc == 0) red_b[ps * CO + kh * KH + k] = accb[k];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red_l[(ps * 2 + kh) * kPwCin + cc * 8 + j] = accL[j];
  __syncthreads();
  float* dst = partial + (size_t)blockIdx.x * pw_partial_floats<CO>();
  for (int o = tid; o < pw_partial_floats<CO>(); o += kPwThreads) {
    float s = 0.f;
    if (o < CO * kPwCin) {
#pragma unroll
      for (int i = 0; i < 16; ++i) s += red[i * CO * kPwCin + o];
    } else if (o < CO * kPwCin + CO) {
#pragma unroll
      for (int i = 0; i < 16; ++i) s += red_b[i * CO + (o - CO * kPwCin)];
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) s += red_l[i * kPwCin + (o - CO * kPwCin - CO)];
    }
    dst[o] = s;
  }
}

// dW / db / dbL (+)= the CTA partials, added in CTA order in double precision
__global__ void pointwise_head_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int CO, int cout,
                                                   float* __restrict__ dw, float* __restrict__ db,
                                                   float* __restrict__ dbl, int accumulate) {
  pdl_prologue();
  const int stride = CO * kPwCin + CO + kPwCin;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < stride; o += gridDim.x * blockDim.x) {
    float* dst = nullptr;
    if (o < CO * kPwCin) {
      if (o / kPwCin < cout) dst = dw + o;
    } else if (o < CO * kPwCin + CO) {
      if (o - CO * kPwCin < cout && db) dst = db + (o - CO * kPwCin);
    } else if (dbl) {
      dst = dbl + (o - CO * kPwCin - CO);
    }
    if (!dst) continue;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += (double)__ldg(partial + (size_t)b * stride + o);
    *dst = accumulate ? *dst + (float)s : (float)s;
  }
}

static int pw_stage(const float* w, const float* bias, int cout, mmr_stream_t stream) {
  float *wd = nullptr, *bd = nullptr;
  MMR_CUDA_CHECK(cudaGetSymbolAddress(reinterpret_cast<void**>(&wd), c_pw_w));
  MMR_CUDA_CHECK(cudaGetSymbolAddress(reinterpret_cast<void**>(&bd), c_pw_b));
  mmr_launch((pointwise_head_stage_kernel), 1, 256, 0, as_stream(stream), w, bias, cout, wd, bd);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}
static int pw_co(int cout) { return cout <= 4 ? 4 : cout <= 8 ? 8 : cout <= 10 ? 10 : 16; }
static int pw_bwd_blocks() { return 2 * num_sms(); }

}  // namespace mmr

using namespace mmr;

#define PW_DISPATCH(co, ...)                                \
  switch (co) {                                             \
    case 4: { constexpr int CO = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int CO = 8; __VA_ARGS__; } break;   \
    case 10: { constexpr int CO = 10; __VA_ARGS__; } break; \
    default: { constexpr int CO = 16; __VA_ARGS__; } break; \
  }

extern "C" int mmr_pointwise_head_fwd(const void* x, const float* w, const float* bias, int N, int H, int W, int Cin,
                                      int Cout, float* logits, mmr_stream_t stream) {
  MMR_REQUIRE(Cin == kPwCin && Cout >= 1 && Cout <= 16, "pointwise head: Cin must be 64 and 1 <= Cout <= 16 (got %d -> %d)", Cin, Cout);
  MMR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "pointwise head: activation must be 16-byte aligned");
  const int64_t hw = (int64_t)H * W, npix = hw * N;
  int64_t blocks = (npix + kPwThreads - 1) / kPwThreads;
  if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
  if (pw_stage(w, bias, Cout, stream)) return -1;
  PW_DISPATCH(pw_co(Cout), (mmr_launch((pointwise_head_fwd_kernel<CO>), (unsigned)blocks, kPwThreads, 0, as_stream(stream),
                                       reinterpret_cast<const __nv_bfloat16*>(x), Cout, npix, hw, logits)));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int64_t mmr_pointwise_head_bwd_workspace_bytes(int Cout) {
  const int co = pw_co(Cout);
  return (int64_t)pw_bwd_blocks() * (co * kPwCin + co + kPwCin) * (int64_t)sizeof(float);
}

extern "C" int mmr_pointwise_head_bwd(const float* dlogits, const void* x, const float* w, int N, int H, int W,
                                      int Cin, int Cout, int relu_mask, void* dx, float* dw, float* dbias,
                                      float* dbias_producer, int accumulate, float* workspace, mmr_stream_t stream) {
  MMR_REQUIRE(Cin == kPwCin && Cout >= 1 && Cout <= 16, "pointwise head: Cin must be 64 and 1 <= Cout <= 16 (got %d -> %d)", Cin, Cout);
  MMR_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
              "pointwise head: activation and gradient must be 16-byte aligned");
  MMR_REQUIRE(workspace != nullptr && dw != nullptr, "pointwise head: workspace and dw are required");
  const int64_t hw = (int64_t)H * W, npix = hw * N;
  const int co = pw_co(Cout);
  int blocks = pw_bwd_blocks();
  const int64_t ntiles = (npix + kPwTile - 1) / kPwTile;
  if (blocks > ntiles) blocks = (int)ntiles;
  if (pw_stage(w, nullptr, Cout, stream)) return -1;
  PW_DISPATCH(co, {
    const size_t smem = sizeof(PwSmem<CO>);
    static bool attr_set = false;
    if (!attr_set) {
      MMR_CUDA_CHECK(cudaFuncSetAttribute(pointwise_head_bwd_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    mmr_launch((pointwise_head_bwd_kernel<CO>), (unsigned)blocks, kPwThreads, smem, as_stream(stream), dlogits,
               reinterpret_cast<const __nv_bfloat16*>(x), Cout, relu_mask, npix, hw,
               reinterpret_cast<__nv_bfloat16*>(dx), workspace);
  });
  MMR_CUDA_CHECK(cudaGetLastError());
  mmr_launch((pointwise_head_bwd_finalize_kernel), 4, 256, 0, as_stream(stream), (const float*)workspace, blocks, co, Cout, dw,
             dbias, dbias_producer, accumulate);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}
