// HBM-bound kernels of the path: layout packing, BatchNorm (+residual +ReLU) forward and
// backward, gradient gathering across the dense skip fan-out (with the 2x2 sum-pool that is
// the backward of nearest-x2 upsampling), and the ResNet stem max-pool.
// All activations are NHWC bf16; every thread moves 16 bytes (8 channels) per access.
#include "common.h"
#include "ptx.cuh"

namespace mmr {

constexpr int kEwThreads = 256;
constexpr int kPdlTailIters = 3;  // streaming kernels trigger the next launch this many grid-stride iterations before their end
constexpr int kMaxContrib = 8;

struct ContribList {
  const __nv_bfloat16* ptr[kMaxContrib];
  int pool2[kMaxContrib];
  int n;
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(w[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// 32-byte global accesses (LDG.256 / STG.256, sm_100+): the streaming BatchNorm passes measured 6.3-6.5 TB/s
// with these against 5.5-5.6 TB/s with 16-byte ones (scripts/probe/ew_stream_probe.cu; cudaMemcpy D2D 5.9).
struct alignas(32) Words8 {
  uint32_t v[8];
};
__device__ __forceinline__ Words8 ld256(const __nv_bfloat16* p) {
  Words8 r;
#ifndef MMR_EW_LD_HINT
#define MMR_EW_LD_HINT ""
#endif
  asm volatile("ld.global.L1::no_allocate" MMR_EW_LD_HINT ".v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                 "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st256(__nv_bfloat16* p, const Words8& r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]),
               "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
               : "memory");
}
// Per-channel constants of the 16-channel passes live in shared memory as [C/16][NA][16] floats with every
// 16-channel block skewed by 4 floats, so the float4 reads of the (up to eight) channel groups a warp touches
// fall into different banks.
template <int NA>
__host__ __device__ constexpr int coef16_stride() { return NA * 16 + 4; }

// Sum of all contributions at pixel (n,y,x), channels [c, c+8).
__device__ __forceinline__ void gather8(const ContribList& cl, int n, int y, int x, int c, int H,
                                        int W, int C, float (&g)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = 0.f;
  for (int i = 0; i < cl.n; ++i) {
    float v[8];
    if (!cl.pool2[i]) {
      load8(cl.ptr[i] + (((size_t)n * H + y) * W + x) * C + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += v[j];
    } else {
      const int H2 = 2 * H, W2 = 2 * W;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          load8(cl.ptr[i] + (((size_t)n * H2 + 2 * y + dy) * W2 + 2 * x + dx) * C + c, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += v[j];
        }
    }
  }
}

static int fill_contribs(ContribList& cl, const MmrContrib* contribs, int n) {
  MMR_REQUIRE(n >= 0 && n <= kMaxContrib, "at most %d gradient contributions, got %d", kMaxContrib, n);
  cl.n = n;
  for (int i = 0; i < kMaxContrib; ++i) {
    cl.ptr[i] = i < n ? reinterpret_cast<const __nv_bfloat16*>(contribs[i].ptr) : nullptr;
    cl.pool2[i] = i < n ? contribs[i].pool2 : 0;
  }
  return 0;
}

static int ew_blocks(int64_t work_items, int cap_mult = 8) {
  int64_t b = (work_items + kEwThreads - 1) / kEwThreads;
  const int64_t cap = (int64_t)num_sms() * cap_mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// The 16-channel (32-byte) passes need C % 16 == 0, 32-byte aligned tensors and constants that fit the default
// 48 KB of dynamic shared memory; everything else takes the 8-channel kernels.
static bool ew16_ok(int C, int n_arrays, const void* a, const void* b, const void* c) {
  if (C % 16 != 0 || (size_t)(C / 16) * (n_arrays * 16 + 4) * sizeof(float) > 48 * 1024) return false;
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 31) == 0;
}

// ------------------------------------------------------------------ packing
// One CTA = 64 consecutive output pixels of one output row: the 3 x 7 x 133 input patch they share
// is staged in shared memory with coalesced row reads (optionally normalised, or taken from uint8 HWC
// frames), then written out as im2col rows [pixel][kpad] in 16-byte pieces.  A per-thread gather
// from global memory ran at 1/8 of the write bandwidth.
constexpr int kStemSeg = 64;
constexpr int kStemTw = 2 * kStemSeg + 5;   // input columns under one segment
constexpr int kStemTwp = kStemTw + 3;       // padded row pitch
template <typename TIn>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const TIn* __restrict__ x, int N, int H, int W, int Ho, int Wo,
                   __nv_bfloat16* __restrict__ out, int kpad, const float* __restrict__ mean,
                   const float* __restrict__ std_) {
  pdl_prologue();
  __shared__ float tile[3 * 7 * kStemTwp];
  __shared__ int lut[256];
  const int ox0 = blockIdx.x * kStemSeg, oy = blockIdx.y, n = blockIdx.z;
  for (int k = threadIdx.x; k < 256; k += blockDim.x) {
    const int c = k / 49, r = k % 49;
    lut[k] = k < 147 ? (c * 7 + r / 7) * kStemTwp + r % 7 : -1;
  }
  for (int idx = threadIdx.x; idx < 3 * 7 * kStemTw; idx += blockDim.x) {
    const int col = idx % kStemTw, rr = idx / kStemTw, ky = rr % 7, c = rr / 7;
    const int iy = 2 * oy + ky - 3, ix = 2 * ox0 + col - 3;
    float val = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      if (sizeof(TIn) == 1)   // uint8 frames, HWC: ToTensor's /255 happens here
        val = (float)x[(((size_t)n * H + iy) * W + ix) * 3 + c] * (1.f / 255.f);
      else
        val = (float)x[(((size_t)n * 3 + c) * H + iy) * W + ix];
      if (mean) val = (val - __ldg(mean + c)) / __ldg(std_ + c);
    }
    tile[(c * 7 + ky) * kStemTwp + col] = val;
  }
  __syncthreads();
  const int groups = kpad / 8;
  const int npx = min(kStemSeg, Wo - ox0);
  const size_t pix0 = ((size_t)n * Ho + oy) * Wo + ox0;
  for (int item = threadIdx.x; item < npx * groups; item += blockDim.x) {
    const int p = item / groups, g = item % groups;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = lut[g * 8 + j];
      v[j] = o >= 0 ? tile[o + 2 * p] : 0.f;
    }
    store8(out + (pix0 + p) * kpad + g * 8, v);
  }
}

// Space-to-depth stem (see mmrseg.h): thread = (block pixel, 8-channel group) of the [N][H/4][W/4][64] tensor.
template <typename TIn>
__global__ void __launch_bounds__(kEwThreads)
stem_s2d_pack_kernel(const TIn* __restrict__ x, int N, int H, int W, __nv_bfloat16* __restrict__ out,
                     const float* __restrict__ mean, const float* __restrict__ std_) {
  pdl_prologue();
  const int Hb = H / 4, Wb = W / 4;
  const uint32_t total = (uint32_t)N * Hb * Wb * 8;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i & 7);
    const uint32_t pix = i >> 3;
    const int X = (int)(pix % (uint32_t)Wb), Y = (int)((pix / (uint32_t)Wb) % (uint32_t)Hb);
    const int n = (int)(pix / ((uint32_t)Wb * Hb));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = g * 8 + j;   // (ry*4 + rx)*3 + c
      float val = 0.f;
      if (ch < 48) {
        const int c = ch % 3, rx = (ch / 3) & 3, ry = ch / 12;
        const int iy = 4 * Y + ry, ix = 4 * X + rx;
        if (sizeof(TIn) == 1)
          val = (float)x[(((size_t)n * H + iy) * W + ix) * 3 + c] * (1.f / 255.f);
        else
          val = (float)x[(((size_t)n * 3 + c) * H + iy) * W + ix];
        if (mean) val = (val - __ldg(mean + c)) / __ldg(std_ + c);
      }
      v[j] = val;
    }
    store8(out + (size_t)pix * 64 + g * 8, v);
  }
}

__global__ void stem_s2d_weights_kernel(const float* __restrict__ w7, int Cout, float* __restrict__ w3) {
  pdl_prologue();
  const int total = 4 * Cout * 64 * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int B = i % 3, A = (i / 3) % 3, ch = (i / 9) % 64, row = i / (9 * 64);
    const int co = row % Cout, q = row / Cout, qy = q >> 1, qx = q & 1;
    float v = 0.f;
    if (ch < 48) {
      const int c = ch % 3, rx = (ch / 3) & 3, ry = ch / 12;
      const int ky = 4 * (A - 1) + ry - 2 * qy + 3, kx = 4 * (B - 1) + rx - 2 * qx + 3;
      if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w7[((co * 3 + c) * 7 + ky) * 7 + kx];
    }
    w3[i] = v;
  }
}

// dW of the 7x7 stem from the gradient of its phase-expanded 3x3 form: every original tap occurs once per
// output phase.  dw7[co][c][ky][kx] (+)= sum over (qy, qx) of dw3[q*Cout + co][(ry*4 + rx)*3 + c][A][B].
__global__ void stem_s2d_wgrad_fold_kernel(const float* __restrict__ dw3, int Cout, float* __restrict__ dw7,
                                           int accumulate) {
  pdl_prologue();
  const int total = Cout * 3 * 49;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kx = i % 7, ky = (i / 7) % 7, c = (i / 49) % 3, co = i / 147;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ty = ky + 2 * (q >> 1) - 3 + 4, tx = kx + 2 * (q & 1) - 3 + 4;   // + 4: floor division of negatives
      const int A = ty / 4, ry = ty % 4, B = tx / 4, rx = tx % 4;
      s += dw3[(((size_t)(q * Cout + co)) * 64 + (ry * 4 + rx) * 3 + c) * 9 + A * 3 + B];
    }
    dw7[i] = accumulate ? dw7[i] + s : s;
  }
}

__global__ void pack_nchw_kernel(const float* __restrict__ x, int N, int C, int H, int W,
                                 __nv_bfloat16* __restrict__ out, int cpad) {
  pdl_prologue();
  const int groups = cpad / 8;
  const int64_t hw = (int64_t)H * W;
  const int64_t total = (int64_t)N * hw * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    // pixel index fastest so that the strided NCHW reads coalesce across the warp
    const int64_t p = i % hw;
    const int g = (int)((i / hw) % groups);
    const int n = (int)(i / (hw * groups));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      v[j] = c < C ? __ldg(x + ((size_t)n * C + c) * hw + p) : 0.f;
    }
    store8(out + ((size_t)n * hw + p) * cpad + g * 8, v);
  }
}

// uint8 HWC frames -> NHWC bf16 padded to cpad channels: x/255, optional (x-mean)/std.
__global__ void pack_u8_kernel(const uint8_t* __restrict__ x, int64_t npix, __nv_bfloat16* __restrict__ out,
                               int cpad, const float* __restrict__ mean, const float* __restrict__ std_) {
  pdl_prologue();
  const int groups = cpad / 8;
  const int64_t total = npix * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const int64_t p = i / groups;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float val = 0.f;
      if (c < 3) {
        val = (float)x[p * 3 + c] * (1.f / 255.f);
        if (mean) val = (val - __ldg(mean + c)) / __ldg(std_ + c);
      }
      v[j] = val;
    }
    store8(out + p * cpad + g * 8, v);
  }
}

__global__ void unpack_nhwc_kernel(const __nv_bfloat16* __restrict__ x, int N, int C, int ldc, int H,
                                   int W, float* __restrict__ out) {
  pdl_prologue();
  const int64_t hw = (int64_t)H * W;
  const int64_t total = (int64_t)N * C * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i % hw;
    const int c = (int)((i / hw) % C);
    const int n = (int)(i / (hw * C));
    out[i] = __bfloat162float(x[((size_t)n * hw + p) * ldc + c]);
  }
}

__global__ void repack_weights_kernel(const float* __restrict__ w, int O, int I, int taps,
                                      __nv_bfloat16* __restrict__ fwd, int ldf,
                                      __nv_bfloat16* __restrict__ dgrad, int ldd, int o_pad) {
  pdl_prologue();
  const int64_t total = (int64_t)O * I * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int ci = (int)((i / taps) % I);
    const int o = (int)(i / ((int64_t)taps * I));
    const __nv_bfloat16 v = __float2bfloat16(w[i]);
    if (fwd) fwd[(size_t)o * ldf + (size_t)t * I + ci] = v;
    if (dgrad) dgrad[(size_t)ci * ldd + (size_t)t * o_pad + o] = v;
  }
}

// ------------------------------------------------------------------ BatchNorm forward
// Thread layout for [P][C] reductions: tpc = C/8 threads span one pixel row, the block's
// 256/tpc thread-rows stride over pixels, two rows per thread and iteration so that every load of
// both rows is in flight before the first use.  Per-thread fp32 partial sums are flushed to double
// every 16 iterations; block partials are written (not atomically added) so results are
// deterministic.  Row indices are 32-bit (P < 2^31); (n, y, x) is only decomposed when a
// contribution is 2x2-pooled, with shifts when H and W are powers of two.
struct RowGeom {
  int H, W, wshift, hshift;  // shifts are -1 when the extent is not a power of two
};
__device__ __forceinline__ void row_to_nyx(const RowGeom& g, uint32_t r, int& n, int& y, int& x) {
  if (g.wshift >= 0 && g.hshift >= 0) {
    x = (int)(r & (uint32_t)(g.W - 1));
    const uint32_t t = r >> g.wshift;
    y = (int)(t & (uint32_t)(g.H - 1));
    n = (int)(t >> g.hshift);
  } else {
    x = (int)(r % (uint32_t)g.W);
    const uint32_t t = r / (uint32_t)g.W;
    y = (int)(t % (uint32_t)g.H);
    n = (int)(t / (uint32_t)g.H);
  }
}

__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void add8(const uint4& u, float (&g)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(w[j]);
    g[2 * j] += f.x;
    g[2 * j + 1] += f.y;
  }
}
__device__ __forceinline__ void cvt8(const uint4& u, float (&g)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(w[j]);
    g[2 * j] = f.x;
    g[2 * j + 1] = f.y;
  }
}

// Any number of contributions (the NC = 0 instantiations: more than four consumers, e.g. the encoder
// features every decoder column reads).  All same-resolution loads of the row are issued first (up to
// eight 16-byte loads in flight), then each 2x2-pooled contribution issues its four loads together.
template <bool POOLED>
__device__ __forceinline__ void gather_row(const ContribList& cl, const RowGeom& geo, uint32_t r, int c, int C,
                                           float (&g)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = 0.f;
  uint4 raw[kMaxContrib];
#pragma unroll
  for (int i = 0; i < kMaxContrib; ++i)
    if (i < cl.n && !(POOLED && cl.pool2[i])) raw[i] = ldg16(cl.ptr[i] + (size_t)r * C + c);
  if (POOLED) {
    int n = 0, y = 0, x = 0;
    row_to_nyx(geo, r, n, y, x);
    const int W2 = 2 * geo.W;
    for (int i = 0; i < cl.n; ++i) {
      if (!cl.pool2[i]) continue;
      const __nv_bfloat16* base = cl.ptr[i] + (((size_t)n * (2 * geo.H) + 2 * y) * W2 + 2 * x) * C + c;
      const uint4 v0 = ldg16(base), v1 = ldg16(base + C), v2 = ldg16(base + (size_t)W2 * C),
                  v3 = ldg16(base + (size_t)W2 * C + C);
      float t[8];
      cvt8(v0, t);
      add8(v1, t);
      float u[8];
      cvt8(v2, u);
      add8(v3, u);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += t[j] + u[j];
    }
  }
#pragma unroll
  for (int i = 0; i < kMaxContrib; ++i)
    if (i < cl.n && !(POOLED && cl.pool2[i])) add8(raw[i], g);
}

// All loads of a row are issued before the first use: the contribution count NC is a template
// parameter (1..4; 0 = any count, serial gather), so the gather unrolls into independent loads.
template <int NC, bool POOLED>
struct RowRaw {
  uint4 v[NC > 0 ? NC : 1][POOLED ? 4 : 1];
  uint4 a, z;
};

template <int MODE, int NC, bool POOLED>
__device__ __forceinline__ void issue_row(const ContribList& cl, const RowGeom& geo, uint32_t r, int c, int C,
                                          const __nv_bfloat16* act, const __nv_bfloat16* z,
                                          RowRaw<NC, POOLED>& raw) {
  int n = 0, y = 0, x = 0;
  if (POOLED) row_to_nyx(geo, r, n, y, x);
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    if (POOLED && cl.pool2[i]) {
      const int W2 = 2 * geo.W;
      const __nv_bfloat16* base = cl.ptr[i] + (((size_t)n * (2 * geo.H) + 2 * y) * W2 + 2 * x) * C + c;
      raw.v[i][0] = ldg16(base);
      raw.v[i][POOLED ? 1 : 0] = ldg16(base + C);
      raw.v[i][POOLED ? 2 : 0] = ldg16(base + (size_t)W2 * C);
      raw.v[i][POOLED ? 3 : 0] = ldg16(base + (size_t)W2 * C + C);
    } else {
      raw.v[i][0] = ldg16(cl.ptr[i] + (size_t)r * C + c);
      if (POOLED) raw.v[i][1] = raw.v[i][2] = raw.v[i][3] = make_uint4(0, 0, 0, 0);
    }
  }
  if (act) raw.a = ldg16(act + (size_t)r * C + c);
  if (MODE == 1 || MODE == 3) raw.z = ldg16(z + (size_t)r * C + c);
}

template <int MODE, int NC, bool POOLED>
__device__ __forceinline__ void finish_row(const ContribList& cl, const RowGeom& geo, uint32_t r, int c, int C,
                                           const __nv_bfloat16* act, const RowRaw<NC, POOLED>& raw,
                                           __nv_bfloat16* gout, float (&f1)[8], float (&f2)[8],
                                           const float (&msc)[8], const float (&msh)[8]) {
  float g[8];
  if (NC == 0) {
    gather_row<POOLED>(cl, geo, r, c, C, g);
  } else {
    // the first piece converts (0 + v is not folded by the compiler: -0.0), the others add
    cvt8(raw.v[0][0], g);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
#pragma unroll
      for (int q = 0; q < (POOLED ? 4 : 1); ++q)
        if (i + q > 0) add8(raw.v[i][q], g);
    }
  }
  float zz[8];
  if (MODE == 1 || MODE == 3) cvt8(raw.z, zz);
  if (MODE == 3) {
    // ReLU mask recomputed from z with the very FMA bn_apply used: the activation is not re-read
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = fmaf(zz[j], msc[j], msh[j]) > 0.f ? g[j] : 0.f;
  } else if (act) {
    float a[8];
    cvt8(raw.a, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = a[j] > 0.f ? g[j] : 0.f;
  }
  if (gout) store8(gout + (size_t)r * C + c, g);   // NULL: the apply pass recomputes g from the contribution
  if (MODE == 1 || MODE == 3) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f1[j] += g[j], f2[j] += g[j] * zz[j];
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) f1[j] += g[j];
  }
}

// Fused mmr_bn_bwd_finalize (MODE 1): block sums are added to slots [8][2][C] and the last CTA
// (ticket) writes dgamma / dbeta / coef, then re-arms slots and ticket.
struct BwdFused {
  const float* gamma;
  float* dgamma;
  float* dbeta;
  float* coef;
  uint32_t* ticket;  // nullptr: not fused, `partial` receives per-block sums
  int accumulate;
  const float* mask_scale;  // MODE 3: ReLU mask = (z * mask_scale + mask_shift > 0)
  const float* mask_shift;
};

template <int MODE, int NC, bool POOLED>  // MODE 0: stats of z ; 1: bn backward reduce ; 2: plain gradient gather ;
                                          // 3: as 1 with the ReLU mask recomputed from z
__global__ void __launch_bounds__(kEwThreads, 2)
reduce_rows_kernel(const __nv_bfloat16* __restrict__ z, uint32_t P, int C, double* __restrict__ partial,
                   ContribList cl, const __nv_bfloat16* __restrict__ act,
                   const float* __restrict__ mean, const float* __restrict__ invstd, RowGeom geo,
                   __nv_bfloat16* __restrict__ gout, BwdFused fz) {
  pdl_prologue();
  // double accumulators live in shared memory ([j][thread], conflict-free) so that two CTAs of
  // 256 threads fit the register file; the loop itself accumulates in fp32 and flushes every 16
  // iterations.  MODE 1 accumulates sum(g) and sum(g*z); sum(g*xhat) = (sum(g*z) - mean*sum(g))*invstd
  // is formed in double when the block partial is written.
  __shared__ double sd1[8 * kEwThreads];
  __shared__ double sd2[8 * kEwThreads];
  const int tpc = C / 8;
  const int rows_per_iter = kEwThreads / tpc;
  const int cg = threadIdx.x % tpc;
  const int rl = threadIdx.x / tpc;
  const int c = cg * 8;
  float f1[8], f2[8], msc[8], msh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    msc[j] = MODE == 3 ? __ldg(fz.mask_scale + c + j) : 0.f;
    msh[j] = MODE == 3 ? __ldg(fz.mask_shift + c + j) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    f1[j] = f2[j] = 0.f;
    sd1[j * kEwThreads + threadIdx.x] = 0.0;
    sd2[j * kEwThreads + threadIdx.x] = 0.0;
  }
  int since_flush = 0;
  const uint32_t stride = (uint32_t)gridDim.x * rows_per_iter;
  // rows in flight per thread: with 2 x 256 threads per SM a row of (NC + 2) 16-byte loads must be
  // multiplied up to ~100 B per thread to cover the HBM latency-bandwidth product (~44 KB per SM)
  constexpr int R = MODE == 0 ? 4 : (POOLED || NC > 4 ? 1 : (NC == 1 ? 4 : (NC == 2 ? 3 : 2)));
  for (uint32_t r0 = (uint32_t)blockIdx.x * rows_per_iter + rl; r0 < P; r0 += R * stride) {
    if (MODE == 0) {
      uint4 raw[R];
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (r0 + k * stride < P) raw[k] = ldg16(z + (size_t)(r0 + k * stride) * C + c);
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (r0 + k * stride < P) {
          float v[8];
          cvt8(raw[k], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) f1[j] += v[j], f2[j] += v[j] * v[j];
        }
    } else {
      RowRaw<NC, POOLED> raw[R];
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (r0 + k * stride < P) issue_row<MODE, NC, POOLED>(cl, geo, r0 + k * stride, c, C, act, z, raw[k]);
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (r0 + k * stride < P)
          finish_row<MODE, NC, POOLED>(cl, geo, r0 + k * stride, c, C, act, raw[k], gout, f1, f2, msc, msh);
    }
    if (++since_flush == 16) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sd1[j * kEwThreads + threadIdx.x] += (double)f1[j];
        sd2[j * kEwThreads + threadIdx.x] += (double)f2[j];
        f1[j] = f2[j] = 0.f;
      }
      since_flush = 0;
    }
  }
  pdl_done();   // the block reduction and the ticket finalisation overlap the next kernel's launch
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sd1[j * kEwThreads + threadIdx.x] += (double)f1[j];
    sd2[j * kEwThreads + threadIdx.x] += (double)f2[j];
  }
  if (partial == nullptr) return;
  __syncthreads();
  // block sums: one thread per channel (not one per channel group) adds the rows_per_iter row lanes in lane order
  for (int t = threadIdx.x; t < C; t += kEwThreads) {
    const int j = t / tpc, g8 = t % tpc, ch = g8 * 8 + j;
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < rows_per_iter; ++k) {
      s1 += sd1[j * kEwThreads + k * tpc + g8];
      s2 += sd2[j * kEwThreads + k * tpc + g8];
    }
    if (MODE == 1 || MODE == 3) s2 = (s2 - (double)__ldg(mean + ch) * s1) * (double)__ldg(invstd + ch);
    if ((MODE == 1 || MODE == 3) && fz.ticket) {
      double* sl = partial + (size_t)(blockIdx.x & 7) * 2 * C;
      atomicAdd(sl + ch, s1);
      atomicAdd(sl + C + ch, s2);
    } else {
      partial[((size_t)blockIdx.x * 2 + 0) * C + ch] = s1;
      partial[((size_t)blockIdx.x * 2 + 1) * C + ch] = s2;
    }
  }
  if ((MODE == 1 || MODE == 3) && fz.ticket) {
    __shared__ uint32_t last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(fz.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int ch = threadIdx.x; ch < C; ch += kEwThreads) {
      double s1 = 0.0, s2 = 0.0;
      for (int sl = 0; sl < 8; ++sl) {
        double* a = partial + (size_t)sl * 2 * C + ch;
        s1 += __ldcg(a);
        s2 += __ldcg(a + C);
        a[0] = 0.0;
        a[C] = 0.0;
      }
      if (fz.dgamma) fz.dgamma[ch] = (fz.accumulate ? fz.dgamma[ch] : 0.f) + (float)s2;
      if (fz.dbeta) fz.dbeta[ch] = (fz.accumulate ? fz.dbeta[ch] : 0.f) + (float)s1;
      const double gi = (double)(fz.gamma ? fz.gamma[ch] : 1.f) * (double)invstd[ch];
      fz.coef[ch] = (float)gi;
      fz.coef[C + ch] = (float)(-gi * s2 / (double)P);
      fz.coef[2 * C + ch] = (float)(-gi * s1 / (double)P);
    }
    if (threadIdx.x == 0) *fz.ticket = 0u;
  }
}

template <int MODE>
static void launch_reduce(int nblk, cudaStream_t st, const __nv_bfloat16* z, uint32_t P, int C, double* partial,
                          const ContribList& cl, const __nv_bfloat16* act, const float* mean, const float* invstd,
                          RowGeom geo, __nv_bfloat16* g, BwdFused fz = BwdFused{}) {
  bool pooled = false;
  for (int i = 0; i < cl.n; ++i) pooled |= cl.pool2[i] != 0;
#define MMR_RR(NC, PL) \
  mmr_launch((reduce_rows_kernel<MODE, NC, PL>), nblk, kEwThreads, 0, st, z, P, C, partial, cl, act, mean, invstd, geo, g, fz)
  if (pooled) {
    switch (cl.n) {
      case 1: MMR_RR(1, true); break;
      case 2: MMR_RR(2, true); break;
      case 3: MMR_RR(3, true); break;
      case 4: MMR_RR(4, true); break;
      default: MMR_RR(0, true); break;
    }
  } else {
    switch (cl.n) {
      case 1: MMR_RR(1, false); break;
      case 2: MMR_RR(2, false); break;
      case 3: MMR_RR(3, false); break;
      case 4: MMR_RR(4, false); break;
      case 5: MMR_RR(5, false); break;   // x_0_0 .. x_0_3 read by five decoder nodes
      case 6: MMR_RR(6, false); break;
      default: MMR_RR(0, false); break;
    }
  }
#undef MMR_RR
}

// Sum partial[b][which][c] over the nblk blocks for 8 consecutive channels per CTA: 256 threads =
// 8 channels x 32 block-lanes, tree-combined in shared memory in a fixed order (deterministic).
__device__ __forceinline__ void reduce_partials8(const double* __restrict__ partial, int nblk, int C,
                                                 int c0, double& s1, double& s2) {
  __shared__ double sh[2][32][8];
  const int j = threadIdx.x & 7, lane = threadIdx.x >> 3;
  const int c = c0 + j;
  double a1 = 0.0, a2 = 0.0;
  if (c < C)
    for (int b = lane; b < nblk; b += 32) {
      a1 += partial[((size_t)b * 2 + 0) * C + c];
      a2 += partial[((size_t)b * 2 + 1) * C + c];
    }
  sh[0][lane][j] = a1;
  sh[1][lane][j] = a2;
  __syncthreads();
  for (int s = 16; s > 0; s >>= 1) {
    if (lane < s) {
      sh[0][lane][j] += sh[0][lane + s][j];
      sh[1][lane][j] += sh[1][lane + s][j];
    }
    __syncthreads();
  }
  s1 = sh[0][0][j];
  s2 = sh[1][0][j];
}

__global__ void __launch_bounds__(256)
bn_finalize_kernel(const double* __restrict__ partial, int nblk, int64_t P, int C,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   float momentum, float* running_mean, float* running_var, int64_t* nbt, float* mean,
                   float* invstd, float* scale, float* shift) {
  pdl_prologue();
  double s1, s2;
  reduce_partials8(partial, nblk, C, blockIdx.x * 8, s1, s2);
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) nbt[0] += 1;
  const int c = blockIdx.x * 8 + (threadIdx.x & 7);
  if ((threadIdx.x >> 3) != 0 || c >= C) return;
  const double m = s1 / (double)P;
  double var = s2 / (double)P - m * m;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean[c] = (float)m;
  invstd[c] = is;
  scale[c] = g * is;
  shift[c] = b - (float)m * g * is;
  if (running_mean) {
    const double unbiased = P > 1 ? var * (double)P / (double)(P - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// Streaming apply: when the grid stride is a multiple of C/8 every thread keeps the same channel
// group for its whole loop (FIXED), so scale / shift are loaded once; two items per iteration.
template <bool FIXED>
__global__ void __launch_bounds__(kEwThreads)
bn_apply_kernel(const __nv_bfloat16* __restrict__ z, int64_t P, int C,
                const float* __restrict__ scale, const float* __restrict__ shift,
                const __nv_bfloat16* __restrict__ residual, int relu,
                __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const uint32_t groups = (uint32_t)C / 8;
  const int64_t total = P * groups;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float sc[8], sh[8];
  auto load_affine = [&](int c) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c));
    const float4 h1 = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
    sc[0] = s0.x, sc[1] = s0.y, sc[2] = s0.z, sc[3] = s0.w, sc[4] = s1.x, sc[5] = s1.y, sc[6] = s1.z, sc[7] = s1.w;
    sh[0] = h0.x, sh[1] = h0.y, sh[2] = h0.z, sh[3] = h0.w, sh[4] = h1.x, sh[5] = h1.y, sh[6] = h1.z, sh[7] = h1.w;
  };
  if (FIXED) load_affine((int)((uint64_t)i % groups) * 8);
  for (; i < total; i += 2 * stride) {
    const int64_t i1 = i + stride;
    const bool two = i1 < total;
    float v0[8], v1[8], r0[8], r1[8];
    load8(z + i * 8, v0);
    if (two) load8(z + i1 * 8, v1);
    if (residual) {
      load8(residual + i * 8, r0);
      if (two) load8(residual + i1 * 8, r1);
    }
    if (!FIXED) load_affine((int)((uint64_t)i % groups) * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v0[j] = v0[j] * sc[j] + sh[j];
    if (residual) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v0[j] += r0[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v0[j] = fmaxf(v0[j], 0.f);
    }
    store8(out + i * 8, v0);
    if (two) {
      if (!FIXED) load_affine((int)((uint64_t)i1 % groups) * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) v1[j] = v1[j] * sc[j] + sh[j];
      if (residual) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v1[j] += r1[j];
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v1[j] = fmaxf(v1[j], 0.f);
      }
      store8(out + i1 * 8, v1);
    }
  }
}

// 16 channels (32 bytes) per thread and step; constants in shared memory (any grid size works).
template <bool RES>
__global__ void __launch_bounds__(kEwThreads)
bn_apply16_kernel(const __nv_bfloat16* __restrict__ z, int64_t total, int C, const float* __restrict__ scale,
                  const float* __restrict__ shift, const __nv_bfloat16* __restrict__ residual, int relu,
                  __nv_bfloat16* __restrict__ out) {
  extern __shared__ float sc[];
  constexpr int S = coef16_stride<2>();
  pdl_prologue();
  for (int t = threadIdx.x; t < C; t += blockDim.x) {
    float* d = sc + (t >> 4) * S + (t & 15);
    d[0] = scale[t];
    d[16] = shift[t];
  }
  __syncthreads();
  const uint32_t groups = (uint32_t)C / 16;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t gi = (uint32_t)(i % groups);
  const uint32_t gstep = (uint32_t)(stride % groups);
  const int64_t i_tail = total - kPdlTailIters * stride;
  for (; i < total; i += stride) {
    if (i >= i_tail) pdl_done();   // last iterations: the next kernel's launch latency hides behind them
    const Words8 zv = ld256(z + i * 16);
    Words8 rv;
    if (RES) rv = ld256(residual + i * 16);
    const float* k = sc + gi * S;
    Words8 o;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = *reinterpret_cast<const float4*>(k + 4 * q), b = *reinterpret_cast<const float4*>(k + 16 + 4 * q);
      const float2 z0 = unpack_bf16x2(zv.v[2 * q]), z1 = unpack_bf16x2(zv.v[2 * q + 1]);
      float v0 = fmaf(z0.x, a.x, b.x), v1 = fmaf(z0.y, a.y, b.y), v2 = fmaf(z1.x, a.z, b.z), v3 = fmaf(z1.y, a.w, b.w);
      if (RES) {
        const float2 r0 = unpack_bf16x2(rv.v[2 * q]), r1 = unpack_bf16x2(rv.v[2 * q + 1]);
        v0 += r0.x, v1 += r0.y, v2 += r1.x, v3 += r1.y;
      }
      if (relu) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f), v2 = fmaxf(v2, 0.f), v3 = fmaxf(v3, 0.f);
      o.v[2 * q] = pack_bf16x2(v0, v1);
      o.v[2 * q + 1] = pack_bf16x2(v2, v3);
    }
    st256(out + i * 16, o);
    gi += gstep;
    if (gi >= groups) gi -= groups;
  }
}

// ------------------------------------------------------------------ BatchNorm backward
// With g = dL/d(bn output) already masked by ReLU, xhat = (z-mean)*invstd, M = N*H*W:
//   dgamma = sum g*xhat,  dbeta = sum g,
//   dz = gamma*invstd * (g - dbeta/M - xhat*dgamma/M) = coefA*g + coefB*xhat + coefC.
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const double* __restrict__ partial, int nblk, int64_t P, int C,
                       const float* __restrict__ gamma, const float* __restrict__ invstd, float* dgamma,
                       float* dbeta, int accumulate, float* coef) {
  pdl_prologue();
  double s1, s2;
  reduce_partials8(partial, nblk, C, blockIdx.x * 8, s1, s2);
  const int c = blockIdx.x * 8 + (threadIdx.x & 7);
  if ((threadIdx.x >> 3) != 0 || c >= C) return;
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s2;
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s1;
  const double gi = (double)(gamma ? gamma[c] : 1.f) * (double)invstd[c];
  coef[c] = (float)gi;
  coef[C + c] = (float)(-gi * s2 / (double)P);
  coef[2 * C + c] = (float)(-gi * s1 / (double)P);
}

// dz = cA*g + cB*xhat + cC with xhat = (z-mean)*invstd, folded per channel into
// dz = cA*g + p*z + q,  p = cB*invstd,  q = cC - cB*mean*invstd  (loaded once per thread when FIXED).
template <bool FIXED>
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ z,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ coef, int64_t P, int C, __nv_bfloat16* __restrict__ dz) {
  pdl_prologue();
  const uint32_t groups = (uint32_t)C / 8;
  const int64_t total = P * groups;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float ca[8], pz[8], q[8];
  auto load_coef = [&](int c) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float is = __ldg(invstd + c + j), cb = __ldg(coef + C + c + j);
      ca[j] = __ldg(coef + c + j);
      pz[j] = cb * is;
      q[j] = __ldg(coef + 2 * C + c + j) - cb * __ldg(mean + c + j) * is;
    }
  };
  if (FIXED) load_coef((int)((uint64_t)i % groups) * 8);
  for (; i < total; i += 2 * stride) {
    const int64_t i1 = i + stride;
    const bool two = i1 < total;
    float g0[8], z0[8], g1[8], z1[8], o[8];
    load8(g + i * 8, g0);
    load8(z + i * 8, z0);
    if (two) {
      load8(g + i1 * 8, g1);
      load8(z + i1 * 8, z1);
    }
    if (!FIXED) load_coef((int)((uint64_t)i % groups) * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = ca[j] * g0[j] + pz[j] * z0[j] + q[j];
    store8(dz + i * 8, o);
    if (two) {
      if (!FIXED) load_coef((int)((uint64_t)i1 % groups) * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = ca[j] * g1[j] + pz[j] * z1[j] + q[j];
      store8(dz + i1 * 8, o);
    }
  }
}

// The same for a unit whose only gradient contribution is one full-resolution tensor dx and whose ReLU
// mask is a function of z (no residual): g = (z*msc + msh > 0) ? dx : 0 is recomputed here, so the
// reduction pass never writes g and this pass reads dx instead (2 B/element less traffic per unit).
template <bool FIXED>
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_apply_masked_kernel(const __nv_bfloat16* __restrict__ dx, const __nv_bfloat16* __restrict__ z,
                           const float* __restrict__ mean, const float* __restrict__ invstd,
                           const float* __restrict__ coef, const float* __restrict__ msc_,
                           const float* __restrict__ msh_, int64_t P, int C, __nv_bfloat16* __restrict__ dz) {
  pdl_prologue();
  const uint32_t groups = (uint32_t)C / 8;
  const int64_t total = P * groups;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float ca[8], pz[8], q[8], msc[8], msh[8];
  auto load_coef = [&](int c) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float is = __ldg(invstd + c + j), cb = __ldg(coef + C + c + j);
      ca[j] = __ldg(coef + c + j);
      pz[j] = cb * is;
      q[j] = __ldg(coef + 2 * C + c + j) - cb * __ldg(mean + c + j) * is;
      msc[j] = __ldg(msc_ + c + j);
      msh[j] = __ldg(msh_ + c + j);
    }
  };
  if (FIXED) load_coef((int)((uint64_t)i % groups) * 8);
  for (; i < total; i += 2 * stride) {
    const int64_t i1 = i + stride;
    const bool two = i1 < total;
    float g0[8], z0[8], g1[8], z1[8], o[8];
    load8(dx + i * 8, g0);
    load8(z + i * 8, z0);
    if (two) {
      load8(dx + i1 * 8, g1);
      load8(z + i1 * 8, z1);
    }
    if (!FIXED) load_coef((int)((uint64_t)i % groups) * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float g = fmaf(z0[j], msc[j], msh[j]) > 0.f ? g0[j] : 0.f;
      o[j] = ca[j] * g + pz[j] * z0[j] + q[j];
    }
    store8(dz + i * 8, o);
    if (two) {
      if (!FIXED) load_coef((int)((uint64_t)i1 % groups) * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g = fmaf(z1[j], msc[j], msh[j]) > 0.f ? g1[j] : 0.f;
        o[j] = ca[j] * g + pz[j] * z1[j] + q[j];
      }
      store8(dz + i1 * 8, o);
    }
  }
}

// Both backward apply passes with 16 channels (32 bytes) per thread and step, constants in shared memory.
template <bool MASKED>
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_apply16_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ z,
                      const float* __restrict__ mean, const float* __restrict__ invstd,
                      const float* __restrict__ coef, const float* __restrict__ msc, const float* __restrict__ msh,
                      int64_t total, int C, __nv_bfloat16* __restrict__ dz) {
  extern __shared__ float sc[];
  constexpr int S = coef16_stride<MASKED ? 5 : 3>();
  pdl_prologue();
  for (int t = threadIdx.x; t < C; t += blockDim.x) {
    float* d = sc + (t >> 4) * S + (t & 15);
    const float is = invstd[t], cb = coef[C + t];
    d[0] = coef[t];
    d[16] = cb * is;
    d[32] = coef[2 * C + t] - cb * mean[t] * is;
    if (MASKED) {
      d[48] = msc[t];
      d[64] = msh[t];
    }
  }
  __syncthreads();
  const uint32_t groups = (uint32_t)C / 16;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t gi = (uint32_t)(i % groups);
  const uint32_t gstep = (uint32_t)(stride % groups);
  const int64_t i_tail = total - kPdlTailIters * stride;
  for (; i < total; i += stride) {
    if (i >= i_tail) pdl_done();
    const Words8 gv = ld256(g + i * 16), zv = ld256(z + i * 16);
    const float* k = sc + gi * S;
    Words8 o;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 ca = *reinterpret_cast<const float4*>(k + 4 * q), pz = *reinterpret_cast<const float4*>(k + 16 + 4 * q),
                   qq = *reinterpret_cast<const float4*>(k + 32 + 4 * q);
      const float2 z0 = unpack_bf16x2(zv.v[2 * q]), z1 = unpack_bf16x2(zv.v[2 * q + 1]);
      float2 g0 = unpack_bf16x2(gv.v[2 * q]), g1 = unpack_bf16x2(gv.v[2 * q + 1]);
      if (MASKED) {
        const float4 ms = *reinterpret_cast<const float4*>(k + 48 + 4 * q), mh = *reinterpret_cast<const float4*>(k + 64 + 4 * q);
        g0.x = fmaf(z0.x, ms.x, mh.x) > 0.f ? g0.x : 0.f;
        g0.y = fmaf(z0.y, ms.y, mh.y) > 0.f ? g0.y : 0.f;
        g1.x = fmaf(z1.x, ms.z, mh.z) > 0.f ? g1.x : 0.f;
        g1.y = fmaf(z1.y, ms.w, mh.w) > 0.f ? g1.y : 0.f;
      }
      o.v[2 * q] = pack_bf16x2(ca.x * g0.x + pz.x * z0.x + qq.x, ca.y * g0.y + pz.y * z0.y + qq.y);
      o.v[2 * q + 1] = pack_bf16x2(ca.z * g1.x + pz.z * z1.x + qq.z, ca.w * g1.y + pz.w * z1.y + qq.w);
    }
    st256(dz + i * 16, o);
    gi += gstep;
    if (gi >= groups) gi -= groups;
  }
}

// ------------------------------------------------------------------ max-pool 3x3 s2 p1
__global__ void __launch_bounds__(kEwThreads)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C,
                   __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ idx) {
  pdl_prologue();
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const uint32_t groups = (uint32_t)C / 8;
  const uint32_t total = (uint32_t)N * Ho * Wo * groups;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % groups) * 8;
    const uint32_t pix = i / groups;
    const int ox = (int)(pix % (uint32_t)Wo), oy = (int)((pix / (uint32_t)Wo) % (uint32_t)Ho);
    const int n = (int)(pix / ((uint32_t)Wo * Ho));
    // all (up to nine) window loads first, then a branch-free scan: a tap outside the image reads as -inf,
    // which can never win, and the argmax starts at the first tap inside the image
    const uint32_t ninf2 = 0xFF80FF80u;   // two bf16 -inf
    uint4 raw[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int iy = 2 * oy + k / 3 - 1, ix = 2 * ox + k % 3 - 1;
      const bool live = iy >= 0 && iy < H && ix >= 0 && ix < W;
      raw[k] = make_uint4(ninf2, ninf2, ninf2, ninf2);
      if (live) raw[k] = ldg16(x + (((size_t)n * H + iy) * W + ix) * C + c);
    }
    float best[8];
    int bi[8];
    const int k0 = (oy == 0 ? 3 : 0) + (ox == 0 ? 1 : 0);
    {
      const uint4 f = oy == 0 ? (ox == 0 ? raw[4] : raw[3]) : (ox == 0 ? raw[1] : raw[0]);
      cvt8(f, best);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) bi[j] = k0;
#pragma unroll
    for (int k = 1; k < 9; ++k) {
      float v[8];
      cvt8(raw[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // torch: first maximum in window scan order wins; NaN propagates.  Taps before k0 are outside the
        // image (-inf: never taken) and tap k0 itself compares equal or is the NaN already held.
        const bool take = v[j] > best[j] || v[j] != v[j];
        best[j] = take ? v[j] : best[j];
        bi[j] = take ? k : bi[j];
      }
    }
    store8(out + (size_t)pix * C + c, best);
    if (idx) {   // NULL in eval mode: the argmax is only needed by the backward pass
      uint2 packed;
      packed.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
      packed.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
      *reinterpret_cast<uint2*>(idx + (size_t)pix * C + c) = packed;
    }
  }
}

// Even H and W, one or two same-resolution contributions: one thread per 2x2 block of input pixels.  The four
// pixels of a block lie in the same four windows (oy in {a, a+1}, ox in {b, b+1}), so each window's argmax bytes
// and gradients are loaded once per block instead of once per pixel (9 window reads per 4 pixels -> 4): the
// per-pixel kernel was bound by those L2 reads (156 us on 16 x 256 x 256 x 64 against ~35 us of HBM traffic).
template <int NC>
__global__ void __launch_bounds__(kEwThreads)
maxpool_bwd_block_kernel(ContribList cl, const uint8_t* __restrict__ idx, int N, int H, int W, int C,
                         __nv_bfloat16* __restrict__ gin) {
  pdl_prologue();
  const int Ho = H / 2, Wo = W / 2;
  const uint32_t groups = (uint32_t)C / 8;
  const uint32_t total = (uint32_t)N * Ho * Wo * groups;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % groups) * 8;
    const uint32_t blk = i / groups;
    const int b = (int)(blk % (uint32_t)Wo), a = (int)((blk / (uint32_t)Wo) % (uint32_t)Ho);
    const int n = (int)(blk / ((uint32_t)Wo * Ho));
    // a window outside the image reads as argmax byte 255 (no tap) and zero gradient: no branches below
    uint2 packed[4];
    uint4 raw[4][NC];
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
      const int oy = a + (wi >> 1), ox = b + (wi & 1);
      packed[wi] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
#pragma unroll
      for (int k = 0; k < NC; ++k) raw[wi][k] = make_uint4(0u, 0u, 0u, 0u);
      if (oy < Ho && ox < Wo) {
        const size_t opix = ((size_t)n * Ho + oy) * Wo + ox;
        packed[wi] = __ldg(reinterpret_cast<const uint2*>(idx + opix * C + c));
#pragma unroll
        for (int k = 0; k < NC; ++k) raw[wi][k] = ldg16(cl.ptr[k] + opix * C + c);
      }
    }
    float acc[4][8];   // [dy * 2 + dx][channel]
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
      float g[8];
      cvt8(raw[wi][0], g);
#pragma unroll
      for (int k = 1; k < NC; ++k) add8(raw[wi][k], g);
      const int wy = wi >> 1, wx = wi & 1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t word = j < 4 ? packed[wi].x : packed[wi].y;
        const int id = (int)((word >> (8 * (j & 3))) & 0xFF);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            if ((wy && !dy) || (wx && !dx)) continue;   // the next window only reaches the odd row / column
            const int tap = (wy ? 0 : 1 + dy) * 3 + (wx ? 0 : 1 + dx);
            acc[dy * 2 + dx][j] += id == tap ? g[j] : 0.f;
          }
      }
    }
    __nv_bfloat16* dst = gin + (((size_t)n * H + 2 * a) * W + 2 * b) * C + c;
    store8(dst, acc[0]);
    store8(dst + C, acc[1]);
    store8(dst + (size_t)W * C, acc[2]);
    store8(dst + (size_t)W * C + C, acc[3]);
  }
}

// ------------------------------------------------------------------ max-pool 2x2 s2 (nn.MaxPool2d(2))
// The in-tree UNet's Down block (SU/UArchModel/unet_parts.py, `Down.maxpool_conv`): disjoint windows,
// floor mode (an odd last row / column is dropped).  idx = position of the first maximum in scan order.
__global__ void maxpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C,
                                    __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ idx) {
  pdl_prologue();
  const int Ho = H / 2, Wo = W / 2;
  const uint32_t groups = (uint32_t)C / 8;
  const uint32_t total = (uint32_t)N * Ho * Wo * groups;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % groups) * 8;
    const uint32_t pix = i / groups;
    const int ox = (int)(pix % (uint32_t)Wo), oy = (int)((pix / (uint32_t)Wo) % (uint32_t)Ho);
    const int n = (int)(pix / ((uint32_t)Wo * Ho));
    const __nv_bfloat16* base = x + (((size_t)n * H + 2 * oy) * W + 2 * ox) * C + c;
    float v[4][8];
    load8(base, v[0]);
    load8(base + C, v[1]);
    load8(base + (size_t)W * C, v[2]);
    load8(base + (size_t)W * C + C, v[3]);
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = v[0][j], bi[j] = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k][j] > best[j] || v[k][j] != v[k][j]) best[j] = v[k][j], bi[j] = k;
    }
    store8(out + (size_t)pix * C + c, best);
    if (idx) {
      uint2 packed;
      packed.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
      packed.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
      *reinterpret_cast<uint2*>(idx + (size_t)pix * C + c) = packed;
    }
  }
}

__global__ void maxpool2_bwd_kernel(ContribList cl, const uint8_t* __restrict__ idx, int N, int H, int W, int C,
                                    __nv_bfloat16* __restrict__ gin) {
  pdl_prologue();
  const int Ho = H / 2, Wo = W / 2;
  const uint32_t groups = (uint32_t)C / 8;
  const uint32_t total = (uint32_t)N * H * W * groups;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % groups) * 8;
    const uint32_t pix = i / groups;
    const int x = (int)(pix % (uint32_t)W), y = (int)((pix / (uint32_t)W) % (uint32_t)H);
    const int n = (int)(pix / ((uint32_t)W * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int oy = y >> 1, ox = x >> 1;
    if (oy < Ho && ox < Wo) {
      const size_t opix = ((size_t)n * Ho + oy) * Wo + ox;
      const uint2 packed = __ldg(reinterpret_cast<const uint2*>(idx + opix * C + c));
      float g[8];
      gather8(cl, n, oy, ox, c, Ho, Wo, C, g);
      const int want = (y & 1) * 2 + (x & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t word = j < 4 ? packed.x : packed.y;
        if ((int)((word >> (8 * (j & 3))) & 0xFF) == want) acc[j] = g[j];
      }
    }
    store8(gin + (size_t)pix * C + c, acc);
  }
}

// gin[n,y,x,c] = sum over the (<=4) windows containing (y,x) whose recorded argmax is (y,x).
// An odd coordinate lies in two windows (taps 0 and 2), an even one in one (tap 1).  Every load of a
// pixel (window argmax bytes + NC contributions per window) is issued before the first use.
template <int NC>
__global__ void __launch_bounds__(kEwThreads)
maxpool_bwd_kernel(ContribList cl, const uint8_t* __restrict__ idx, int N, int H, int W, int C,
                   __nv_bfloat16* __restrict__ gin) {
  pdl_prologue();
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const uint32_t groups = (uint32_t)C / 8;
  const uint32_t total = (uint32_t)N * H * W * groups;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % groups) * 8;
    const uint32_t pix = i / groups;
    const int x = (int)(pix % (uint32_t)W), y = (int)((pix / (uint32_t)W) % (uint32_t)H);
    const int n = (int)(pix / ((uint32_t)W * H));
    // window candidates per axis: (output coordinate, tap)
    int oys[2], kys[2], oxs[2], kxs[2], ny = 0, nx = 0;
    if (y & 1) {
      if ((y + 1) / 2 < Ho) oys[ny] = (y + 1) / 2, kys[ny++] = 0;
      oys[ny] = (y - 1) / 2, kys[ny++] = 2;
    } else if (y / 2 < Ho) {
      oys[ny] = y / 2, kys[ny++] = 1;
    }
    if (x & 1) {
      if ((x + 1) / 2 < Wo) oxs[nx] = (x + 1) / 2, kxs[nx++] = 0;
      oxs[nx] = (x - 1) / 2, kxs[nx++] = 2;
    } else if (x / 2 < Wo) {
      oxs[nx] = x / 2, kxs[nx++] = 1;
    }
    uint2 packed[4];
    uint4 raw[4][NC > 0 ? NC : 1];
    int want[4];
    bool live[4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int wi = a * 2 + b;
        live[wi] = a < ny && b < nx;
        if (live[wi]) {
          const size_t opix = ((size_t)n * Ho + oys[a]) * Wo + oxs[b];
          want[wi] = kys[a] * 3 + kxs[b];
          packed[wi] = __ldg(reinterpret_cast<const uint2*>(idx + opix * C + c));
#pragma unroll
          for (int k = 0; k < NC; ++k) raw[wi][k] = __ldg(reinterpret_cast<const uint4*>(cl.ptr[k] + opix * C + c));
        }
      }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
      if (!live[wi]) continue;
      float g[8];
      if (NC > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) add8(raw[wi][k], g);
      } else {
        const int a = wi >> 1, b = wi & 1;
        gather8(cl, n, oys[a], oxs[b], c, Ho, Wo, C, g);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t word = j < 4 ? packed[wi].x : packed[wi].y;
        const int id = (int)((word >> (8 * (j & 3))) & 0xFF);
        if (id == want[wi]) acc[j] += g[j];
      }
    }
    store8(gin + (size_t)pix * C + c, acc);
  }
}

// ------------------------------------------------------------------ bilinear x2 (align_corners)
// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) of the reference's ResNetUNet
// (SU/UArchModel/resnet_unet.py:195): src = dst * (in-1)/(out-1) in fp32 as ATen computes it,
// i0 = floor(src), i1 = i0 + (i0 < in-1), lambda = src - i0.
struct Lerp {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lerp lerp_coord(int dst, int in, float scale) {
  Lerp l;
  const float src = scale * (float)dst;
  l.i0 = (int)src;
  if (l.i0 > in - 1) l.i0 = in - 1;
  l.i1 = l.i0 + (l.i0 < in - 1 ? 1 : 0);
  l.w1 = src - (float)l.i0;
  l.w0 = 1.f - l.w1;
  return l;
}

// Forward: a thread owns one (output column, 8-channel group) and walks down a strip of output rows.  It keeps
// the horizontally interpolated values of the two source rows of the stencil in registers; moving one output
// row down advances the stencil by at most one source row, so a strip of `rows` output rows reads
// rows / 2 + 1 source rows (two 16-byte loads each) instead of four loads per output element, and all index
// arithmetic but one 32-bit division per thread is block-uniform.  A row of NHWC is contiguous in
// (x, channel), so consecutive threads store consecutive 16-byte pieces.
__global__ void __launch_bounds__(kEwThreads)
upsample_bilinear2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, int H, int W, int C, int rows,
                               __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const int H2 = 2 * H, W2 = 2 * W;
  const uint32_t groups = (uint32_t)C / 8;
  const uint32_t j = blockIdx.x * kEwThreads + threadIdx.x;      // piece of the output row
  if (j >= (uint32_t)W2 * groups) return;
  const float sy = H2 > 1 ? (float)(H - 1) / (float)(H2 - 1) : 0.f;
  const float sx = W2 > 1 ? (float)(W - 1) / (float)(W2 - 1) : 0.f;
  const int ox = (int)(j / groups), c = (int)(j % groups) * 8;
  const Lerp lx = lerp_coord(ox, W, sx);
  const int n = blockIdx.z;
  const __nv_bfloat16* src0 = x + (size_t)n * H * W * C + (size_t)lx.i0 * C + c;
  const __nv_bfloat16* src1 = x + (size_t)n * H * W * C + (size_t)lx.i1 * C + c;
  __nv_bfloat16* dst = out + (size_t)n * H2 * W2 * C + (size_t)j * 8;
  const size_t in_row = (size_t)W * C, out_row = (size_t)W2 * C;
  const int oy0 = blockIdx.y * rows, oy1 = min(H2, oy0 + rows);
  int rowA = -1, rowB = -1;      // source rows held in hA / hB
  float hA[8], hB[8];
  auto hrow = [&](int r, float (&h)[8]) {
    float a[8], b[8];
    load8(src0 + (size_t)r * in_row, a);
    load8(src1 + (size_t)r * in_row, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = lx.w0 * a[k] + lx.w1 * b[k];
  };
  for (int oy = oy0; oy < oy1; ++oy) {
    const Lerp ly = lerp_coord(oy, H, sy);
    if (ly.i0 != rowA) {
      if (ly.i0 == rowB) {
#pragma unroll
        for (int k = 0; k < 8; ++k) hA[k] = hB[k];
      } else {
        hrow(ly.i0, hA);
      }
      rowA = ly.i0;
    }
    if (ly.i1 != rowB) {
      if (ly.i1 == rowA) {
#pragma unroll
        for (int k = 0; k < 8; ++k) hB[k] = hA[k];
      } else {
        hrow(ly.i1, hB);
      }
      rowB = ly.i1;
    }
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = ly.w0 * hA[k] + ly.w1 * hB[k];
    store8(dst + (size_t)oy * out_row, o);
  }
}

// Adjoint, as a gather (deterministic): input pixel (yi, xi) collects every output pixel whose
// interpolation stencil contains it.  With a scale of about 1/2 those lie in [2*i-2, 2*i+3].
__device__ __forceinline__ void lerp_adjoint_weights(int i, int in, float scale, int& first, float (&w)[6]) {
  first = 2 * i - 2;
  const int out = 2 * in;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int d = first + k;
    float wk = 0.f;
    if (d >= 0 && d < out) {
      const Lerp l = lerp_coord(d, in, scale);
      if (l.i0 == i) wk += l.w0;
      if (l.i1 == i) wk += l.w1;  // i0 == i1 on the last row/column: both weights land here
    }
    w[k] = wk;
  }
}

// The adjoint is separable like the forward.  A thread owns one (input column, 8-channel group) and a strip of
// input rows: it walks down the output rows that touch the strip, takes the horizontal adjoint of each
// (its <= 5 taps of that row, weights fixed per thread) and adds it, times the row's two vertical weights,
// to the two input rows of that row's stencil.  Those advance monotonically, so two register accumulators
// are enough: about 9 loads per input element instead of 16, the vertical weights are block-uniform, and
// each gradient row is read by one strip (plus the rows shared with its neighbours).
// SINGLE: one full-resolution contribution (the consumer conv's data gradient), read straight from its pointer.
template <bool SINGLE>
__global__ void __launch_bounds__(kEwThreads)
upsample_bilinear2x_bwd_kernel(ContribList cl, int H, int W, int C, int rows, __nv_bfloat16* __restrict__ gin) {
  pdl_prologue();
  const int H2 = 2 * H, W2 = 2 * W;
  const uint32_t groups = (uint32_t)C / 8;
  const uint32_t j = blockIdx.x * kEwThreads + threadIdx.x;      // piece of the input row
  if (j >= (uint32_t)W * groups) return;
  const float sy = H2 > 1 ? (float)(H - 1) / (float)(H2 - 1) : 0.f;
  const float sx = W2 > 1 ? (float)(W - 1) / (float)(W2 - 1) : 0.f;
  const int xi = (int)(j / groups), c = (int)(j % groups) * 8;
  const int n = blockIdx.z;
  int fx;
  float wx[6];
  lerp_adjoint_weights(xi, W, sx, fx, wx);
  const int y0 = blockIdx.y * rows, y1 = min(H, y0 + rows);
  const int d0 = max(0, 2 * y0 - 3), d1 = min(H2 - 1, 2 * (y1 - 1) + 4);
  __nv_bfloat16* dst = gin + (size_t)n * H * W * C + (size_t)j * 8;
  const size_t in_row = (size_t)W * C;
  // scalars, not an array: the six weights stay in registers (the array went to local memory: 6 LDL per row)
  const float wx0 = wx[0], wx1 = wx[1], wx2 = wx[2], wx3 = wx[3], wx4 = wx[4], wx5 = wx[5];
  const size_t out_row = (size_t)W2 * C;
  // SINGLE: the support of a column is at most five consecutive taps of the six-tap window (an open interval
  // shorter than 5): they start at tap 0 or tap 1
  const int t0 = wx0 != 0.f ? 0 : 1;
  const float tw[5] = {t0 ? wx1 : wx0, t0 ? wx2 : wx1, t0 ? wx3 : wx2, t0 ? wx4 : wx3, t0 ? wx5 : wx4};
  int tap_off[5];      // element offset of each tap's (clamped) column inside a gradient row
#pragma unroll
  for (int k = 0; k < 5; ++k) tap_off[k] = min(max(fx + t0 + k, 0), W2 - 1) * C;
  const __nv_bfloat16* grow = cl.ptr[0] + (size_t)n * H2 * out_row + c;      // row 0 of this image's gradient
  int cur = lerp_coord(d0, H, sy).i0;      // accA belongs to input row cur, accB to cur + 1
  float accA[8], accB[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) accA[k] = accB[k] = 0.f;
  for (int d = d0; d <= d1; ++d) {
    const Lerp ly = lerp_coord(d, H, sy);
    while (ly.i0 > cur) {        // the stencil moved down: row cur is complete
      if (cur >= y0 && cur < y1) store8(dst + (size_t)cur * in_row, accA);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        accA[k] = accB[k];
        accB[k] = 0.f;
      }
      ++cur;
    }
    float h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = 0.f;
    if (SINGLE) {
      // all five loads of the row are issued before the first use (a branch per tap serialised them: one load in
      // flight per thread, 2.8 TB/s); a tap outside the column's support reads a clamped column with weight 0
      const __nv_bfloat16* rowp = grow + (size_t)d * out_row;
      uint4 raw[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) raw[k] = __ldg(reinterpret_cast<const uint4*>(rowp + tap_off[k]));
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const uint32_t w[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16x2(w[q]);
          h[2 * q] += tw[k] * f.x;
          h[2 * q + 1] += tw[k] * f.y;
        }
      }
    } else {
#define MMR_BILINEAR_TAP(K, WK)                                              \
      if (WK != 0.f) { /* taps outside the image or outside this column's support carry weight 0 */ \
        float g[8];                                                          \
        gather8(cl, n, d, fx + (K), c, H2, W2, C, g);                        \
        _Pragma("unroll") for (int q = 0; q < 8; ++q) h[q] += WK * g[q];     \
      }
      MMR_BILINEAR_TAP(0, wx0)
      MMR_BILINEAR_TAP(1, wx1)
      MMR_BILINEAR_TAP(2, wx2)
      MMR_BILINEAR_TAP(3, wx3)
      MMR_BILINEAR_TAP(4, wx4)
      MMR_BILINEAR_TAP(5, wx5)
#undef MMR_BILINEAR_TAP
    }
    const float wa = ly.i1 == ly.i0 ? ly.w0 + ly.w1 : ly.w0;      // last row: both weights land on it
    const float wb = ly.i1 == ly.i0 ? 0.f : ly.w1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      accA[k] += wa * h[k];
      accB[k] += wb * h[k];
    }
  }
  if (cur >= y0 && cur < y1) store8(dst + (size_t)cur * in_row, accA);
  if (cur + 1 >= y0 && cur + 1 < y1) store8(dst + (size_t)(cur + 1) * in_row, accB);
}

// ------------------------------------------------------------------ deep-supervision logits
// fp32 NCHW logits of an auxiliary head at 1/f resolution -> full resolution (nearest), and the
// adjoint (f x f sum-pool) for the gradient.  The reference has no deep supervision;
// the definition is oracle/unetpp.py::DeepSupervisionUnetPlusPlus.
__global__ void upsample_nearest_f32_kernel(const float* __restrict__ in, int64_t planes, int h, int w, int f,
                                            float* __restrict__ out) {
  pdl_prologue();
  const int H = h * f, W = w * f;
  const int64_t total = planes * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int64_t pl = i / ((int64_t)W * H);
    out[i] = __ldg(in + (pl * h + y / f) * w + x / f);
  }
}
// Four output pixels per thread (one 16-byte store; f = 2: two source values, f % 4 == 0: one), 32-bit index
// arithmetic: the per-element kernel above ran at 0.65 TB/s on the 67 MB logits of config 4's auxiliary heads.
template <int F2>   // 1: f == 2, 0: f % 4 == 0
__global__ void __launch_bounds__(kEwThreads)
upsample_nearest_f32_vec4_kernel(const float* __restrict__ in, uint32_t rows, int h, int w, int f,
                                 float* __restrict__ out) {
  pdl_prologue();
  const uint32_t H = (uint32_t)h * f, wv = (uint32_t)w * f / 4;      // float4 pieces per output row
  const uint32_t total = rows * wv;
  for (uint32_t i = blockIdx.x * kEwThreads + threadIdx.x; i < total; i += gridDim.x * kEwThreads) {
    const uint32_t row = i / wv, xv = i - row * wv;
    const uint32_t pl = row / H, y = row - pl * H;
    const float* src = in + ((size_t)pl * h + y / (uint32_t)f) * w;
    float4 o;
    if (F2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(src) + xv);
      o = make_float4(v.x, v.x, v.y, v.y);
    } else {
      const float v = __ldg(src + (xv * 4) / (uint32_t)f);
      o = make_float4(v, v, v, v);
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}
__global__ void sumpool_f32_kernel(const float* __restrict__ in, int64_t planes, int h, int w, int f,
                                   float* __restrict__ out) {
  pdl_prologue();
  const int H = h * f, W = w * f;
  const int64_t total = planes * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const int64_t pl = i / ((int64_t)w * h);
    const float* src = in + (pl * H + (int64_t)y * f) * W + (int64_t)x * f;
    float s = 0.f;
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) s += __ldg(src + (int64_t)dy * W + dx);
    out[i] = s;
  }
}

// ------------------------------------------------------------------ head gradient prep
// dlogits fp32 NCHW -> bf16 NHWC (cpad channels, zero padded); per-class sums -> dbias.
__global__ void head_grad_prep_kernel(const float* __restrict__ dl, int N, int C, int H, int W,
                                      __nv_bfloat16* __restrict__ out, int cpad,
                                      double* __restrict__ partial) {
  pdl_prologue();
  const int64_t hw = (int64_t)H * W;
  const int64_t total = (int64_t)N * hw;
  double acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i % hw;
    const int n = (int)(i / hw);
    for (int g = 0; g < cpad / 8; ++g) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        v[j] = c < C ? __ldg(dl + ((size_t)n * C + c) * hw + p) : 0.f;
        if (g < 2) acc[(g * 8 + j) & 15] += v[j];
      }
      store8(out + (size_t)i * cpad + g * 8, v);
    }
  }
  if (partial == nullptr) return;
  __shared__ double sh[kEwThreads];
  for (int j = 0; j < 16 && j < C; ++j) {
    sh[threadIdx.x] = acc[j];
    __syncthreads();
    for (int s = kEwThreads / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * 16 + j] = sh[0];
    __syncthreads();
  }
}

// The same for the usual head shape (cpad == 16, H*W a multiple of 4): a thread converts four consecutive pixels
// per step -- one float4 load per class, two steps in flight, one 32-byte store per pixel -- instead of one
// 4-byte load per class and pixel (136 -> 40 us on 16 x 2 x 512 x 512).
template <int CM>   // classes rounded up to 4 or 16
__global__ void __launch_bounds__(kEwThreads)
head_grad_prep4_kernel(const float* __restrict__ dl, int N, int C, int64_t hw, __nv_bfloat16* __restrict__ out,
                       double* __restrict__ partial) {
  pdl_prologue();
  const int64_t quads = hw / 4, total = (int64_t)N * quads;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc[CM];
#pragma unroll
  for (int j = 0; j < CM; ++j) acc[j] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    float4 v[2][CM];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t iu = i + u * stride;
      if (iu >= total) continue;
      const int n = (int)(iu / quads);
      const float4* src = reinterpret_cast<const float4*>(dl + (size_t)n * C * hw) + (iu - (int64_t)n * quads);
#pragma unroll
      for (int c = 0; c < CM; ++c) v[u][c] = c < C ? __ldg(src + (size_t)c * quads) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t iu = i + u * stride;
      if (iu >= total) continue;
#pragma unroll
      for (int c = 0; c < CM; ++c) acc[c] += (double)((v[u][c].x + v[u][c].y) + (v[u][c].z + v[u][c].w));
      Words8 o[4];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (2 * q < CM) {
          const float4 a = v[u][2 * q < CM ? 2 * q : 0], b = v[u][2 * q + 1 < CM ? 2 * q + 1 : 0];
          o[0].v[q] = pack_bf16x2(a.x, b.x);
          o[1].v[q] = pack_bf16x2(a.y, b.y);
          o[2].v[q] = pack_bf16x2(a.z, b.z);
          o[3].v[q] = pack_bf16x2(a.w, b.w);
        } else {
          o[0].v[q] = o[1].v[q] = o[2].v[q] = o[3].v[q] = 0u;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) st256(out + ((size_t)iu * 4 + k) * 16, o[k]);
    }
  }
  if (partial == nullptr) return;
  __shared__ double sh[kEwThreads];
  for (int j = 0; j < CM && j < C; ++j) {
    sh[threadIdx.x] = acc[j];
    __syncthreads();
    for (int s = kEwThreads / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * 16 + j] = sh[0];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
head_bias_finalize_kernel(const double* __restrict__ partial, int nblk, int C, float* dbias, int accumulate) {
  pdl_prologue();
  // 16 channels x 16 block-lanes, combined in lane order through shared memory (deterministic)
  __shared__ double sh[16][17];
  const int c = threadIdx.x & 15, lane = threadIdx.x >> 4;
  double s = 0.0;
  if (c < C)
    for (int b = lane; b < nblk; b += 16) s += partial[(size_t)b * 16 + c];
  sh[lane][c] = s;
  __syncthreads();
  if (lane == 0 && c < C) {
    for (int l = 1; l < 16; ++l) s += sh[l][c];
    dbias[c] = (accumulate ? dbias[c] : 0.f) + (float)s;
  }
}

static double* g_head_ws = nullptr;  // small persistent workspace for head_grad_prep partials
static const int kHeadBlocks = 592;

}  // namespace mmr

using namespace mmr;

extern "C" int mmr_stem_im2col(const float* x, int N, int H, int W, void* out, int kpad,
                               const float* mean, const float* std_, mmr_stream_t stream) {
  MMR_REQUIRE(kpad % 8 == 0 && kpad >= 152, "kpad must be a multiple of 8 and >= 152, got %d", kpad);
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  MMR_REQUIRE(kpad <= 256 && Ho <= 65535 && N <= 65535, "stem_im2col: kpad <= 256, Ho and N <= 65535");
  dim3 grid((Wo + kStemSeg - 1) / kStemSeg, Ho, N);
  mmr_launch((stem_im2col_kernel<float>), grid, 256, 0, as_stream(stream), x, N, H, W, Ho, Wo, reinterpret_cast<__nv_bfloat16*>(out), kpad, mean, std_);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_stem_im2col_u8(const uint8_t* x_nhwc, int N, int H, int W, void* out, int kpad,
                                  const float* mean, const float* std_, mmr_stream_t stream) {
  MMR_REQUIRE(kpad % 8 == 0 && kpad >= 152 && kpad <= 256, "kpad must be a multiple of 8 in [152, 256], got %d", kpad);
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  MMR_REQUIRE(Ho <= 65535 && N <= 65535, "stem_im2col: Ho and N <= 65535");
  dim3 grid((Wo + kStemSeg - 1) / kStemSeg, Ho, N);
  mmr_launch((stem_im2col_kernel<uint8_t>), grid, 256, 0, as_stream(stream), x_nhwc, N, H, W, Ho, Wo, reinterpret_cast<__nv_bfloat16*>(out), kpad, mean, std_);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_stem_s2d_pack(const void* x, int is_u8, int N, int H, int W, void* out, const float* mean,
                                 const float* std_, mmr_stream_t stream) {
  MMR_REQUIRE(x && out && H % 4 == 0 && W % 4 == 0, "stem_s2d_pack: H and W must be multiples of 4");
  MMR_REQUIRE((mean == nullptr) == (std_ == nullptr), "pass both mean and std or neither");
  const int64_t total = (int64_t)N * (H / 4) * (W / 4) * 8;
  MMR_REQUIRE(total < ((int64_t)1 << 31), "stem_s2d_pack: tensor too large for 32-bit indexing");
  if (is_u8)
    mmr_launch((stem_s2d_pack_kernel<uint8_t>), ew_blocks(total, 32), kEwThreads, 0, as_stream(stream),
               reinterpret_cast<const uint8_t*>(x), N, H, W, reinterpret_cast<__nv_bfloat16*>(out), mean, std_);
  else
    mmr_launch((stem_s2d_pack_kernel<float>), ew_blocks(total, 32), kEwThreads, 0, as_stream(stream),
               reinterpret_cast<const float*>(x), N, H, W, reinterpret_cast<__nv_bfloat16*>(out), mean, std_);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_stem_s2d_weights(const float* w7, int Cout, float* w3_oihw, mmr_stream_t stream) {
  MMR_REQUIRE(w7 && w3_oihw && Cout >= 1, "bad argument");
  const int total = 4 * Cout * 64 * 9;
  mmr_launch((stem_s2d_weights_kernel), (total + 255) / 256, 256, 0, as_stream(stream), w7, Cout, w3_oihw);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_stem_s2d_wgrad_fold(const float* dw3_oihw, int Cout, float* dw7, int accumulate,
                                       mmr_stream_t stream) {
  MMR_REQUIRE(dw3_oihw && dw7 && Cout >= 1, "bad argument");
  const int total = Cout * 147;
  mmr_launch((stem_s2d_wgrad_fold_kernel), (total + 255) / 256, 256, 0, as_stream(stream), dw3_oihw, Cout, dw7,
             accumulate);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_pack_nchw_f32_to_nhwc_bf16(const float* x, int N, int C, int H, int W, void* out,
                                              int cpad, mmr_stream_t stream) {
  MMR_REQUIRE(cpad % 8 == 0 && cpad >= C, "cpad must be a multiple of 8 and >= C");
  const int64_t total = (int64_t)N * H * W * (cpad / 8);
  mmr_launch((pack_nchw_kernel), ew_blocks(total, 16), kEwThreads, 0, as_stream(stream), x, N, C, H, W, reinterpret_cast<__nv_bfloat16*>(out), cpad);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_pack_nhwc_u8_to_nhwc_bf16(const uint8_t* x, int N, int H, int W, void* out, int cpad,
                                             const float* mean, const float* std_, mmr_stream_t stream) {
  MMR_REQUIRE(cpad % 8 == 0 && cpad >= 8, "cpad must be a multiple of 8");
  MMR_REQUIRE((mean == nullptr) == (std_ == nullptr), "pass both mean and std or neither");
  const int64_t npix = (int64_t)N * H * W;
  mmr_launch((pack_u8_kernel), ew_blocks(npix * (cpad / 8), 16), kEwThreads, 0, as_stream(stream), x, npix, reinterpret_cast<__nv_bfloat16*>(out), cpad, mean, std_);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_unpack_nhwc_bf16_to_nchw_f32(const void* x, int N, int C, int ldc, int H, int W,
                                                float* out, mmr_stream_t stream) {
  const int64_t total = (int64_t)N * C * H * W;
  mmr_launch((unpack_nhwc_kernel), ew_blocks(total, 16), kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(x), N, C, ldc, H, W, out);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_repack_weights(const float* w, int O, int I, int taps, void* fwd, int ldf,
                                  void* dgrad, int ldd, int o_pad, mmr_stream_t stream) {
  MMR_REQUIRE(!fwd || ldf >= taps * I, "forward row stride too small");
  MMR_REQUIRE(!dgrad || (o_pad >= O && ldd >= taps * o_pad), "dgrad row stride too small");
  const int64_t total = (int64_t)O * I * taps;
  mmr_launch((repack_weights_kernel), ew_blocks(total, 4), kEwThreads, 0, as_stream(stream), w, O, I, taps, reinterpret_cast<__nv_bfloat16*>(fwd), ldf,
      reinterpret_cast<__nv_bfloat16*>(dgrad), ldd, o_pad);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static RowGeom make_geom(int H, int W) {
  auto lg = [](int v) {
    int s = 0;
    while ((1 << s) < v) ++s;
    return (1 << s) == v ? s : -1;
  };
  return RowGeom{H, W, lg(W), lg(H)};
}
static int check_rows_layout(int C) {
  MMR_REQUIRE(C % 8 == 0 && C >= 8 && C <= 2048 && (kEwThreads % (C / 8)) == 0,
              "channel count %d unsupported by the row-reduction layout (need C/8 | 256)", C);
  return 0;
}

extern "C" int mmr_bn_stats(const void* z, int64_t P, int C, double* partial, int nblk,
                            mmr_stream_t stream) {
  if (check_rows_layout(C)) return -1;
  ContribList cl;
  fill_contribs(cl, nullptr, 0);
  MMR_REQUIRE(P < ((int64_t)1 << 31), "row count must be below 2^31");
  mmr_launch((reduce_rows_kernel<0, 0, false>), nblk, kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(z), (uint32_t)P, C, partial, cl, nullptr, nullptr, nullptr,
      make_geom(1, 1), nullptr, BwdFused{});
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_bn_finalize(const double* partial, int nblk, int64_t P, int C, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean,
                               float* running_var, int64_t* nbt, float* mean, float* invstd,
                               float* scale, float* shift, mmr_stream_t stream) {
  mmr_launch((bn_finalize_kernel), (C + 7) / 8, 256, 0, as_stream(stream), partial, nblk, P, C, gamma, beta, eps, momentum, running_mean, running_var, nbt, mean, invstd,
      scale, shift);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_bn_apply(const void* z, int64_t P, int C, const float* scale, const float* shift,
                            const void* residual, int relu, void* out, mmr_stream_t stream) {
  MMR_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
  if (ew16_ok(C, 2, z, residual, out)) {
    const int64_t total = P * (C / 16);
    const size_t smem = (size_t)(C / 16) * coef16_stride<2>() * sizeof(float);
    const __nv_bfloat16 *zp = reinterpret_cast<const __nv_bfloat16*>(z), *rp = reinterpret_cast<const __nv_bfloat16*>(residual);
    if (residual)
      mmr_launch((bn_apply16_kernel<true>), ew_blocks(total, 32), kEwThreads, smem, as_stream(stream), zp, total, C,
                 scale, shift, rp, relu, reinterpret_cast<__nv_bfloat16*>(out));
    else
      mmr_launch((bn_apply16_kernel<false>), ew_blocks(total, 32), kEwThreads, smem, as_stream(stream), zp, total, C,
                 scale, shift, rp, relu, reinterpret_cast<__nv_bfloat16*>(out));
    MMR_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  const int blocks = ew_blocks((P * (C / 8) + 1) / 2, 16);
  if (((int64_t)blocks * kEwThreads) % (C / 8) == 0)
    mmr_launch((bn_apply_kernel<true>), blocks, kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(z), P, C, scale, shift,
        reinterpret_cast<const __nv_bfloat16*>(residual), relu, reinterpret_cast<__nv_bfloat16*>(out));
  else
    mmr_launch((bn_apply_kernel<false>), blocks, kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(z), P, C, scale, shift,
        reinterpret_cast<const __nv_bfloat16*>(residual), relu, reinterpret_cast<__nv_bfloat16*>(out));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_bn_bwd_reduce(const MmrContrib* contribs, int ncontrib, const void* act,
                                 const void* z, const float* mean, const float* invstd, int N, int H,
                                 int W, int C, void* g, double* partial, int nblk,
                                 mmr_stream_t stream) {
  if (check_rows_layout(C)) return -1;
  ContribList cl;
  if (fill_contribs(cl, contribs, ncontrib)) return -1;
  const int64_t P = (int64_t)N * H * W;
  MMR_REQUIRE(P < ((int64_t)1 << 31), "row count must be below 2^31");
  launch_reduce<1>(nblk, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(z), (uint32_t)P, C, partial, cl,
                   reinterpret_cast<const __nv_bfloat16*>(act), mean, invstd, make_geom(H, W),
                   reinterpret_cast<__nv_bfloat16*>(g));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_bn_bwd_finalize(const double* partial, int nblk, int64_t P, int C,
                                   const float* gamma, const float* invstd, float* dgamma,
                                   float* dbeta, int accumulate, float* coef, mmr_stream_t stream) {
  mmr_launch((bn_bwd_finalize_kernel), (C + 7) / 8, 256, 0, as_stream(stream), partial, nblk, P, C, gamma, invstd, dgamma, dbeta, accumulate, coef);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_bn_bwd_apply(const void* g, const void* z, const float* mean, const float* invstd,
                                const float* coef, int64_t P, int C, void* dz, mmr_stream_t stream) {
  MMR_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
  if (ew16_ok(C, 3, g, z, dz)) {
    const int64_t total = P * (C / 16);
    mmr_launch((bn_bwd_apply16_kernel<false>), ew_blocks(total, 32), kEwThreads,
               (size_t)(C / 16) * coef16_stride<3>() * sizeof(float), as_stream(stream),
               reinterpret_cast<const __nv_bfloat16*>(g), reinterpret_cast<const __nv_bfloat16*>(z), mean, invstd, coef,
               (const float*)nullptr, (const float*)nullptr, total, C, reinterpret_cast<__nv_bfloat16*>(dz));
    MMR_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  const int blocks = ew_blocks((P * (C / 8) + 1) / 2, 16);
  if (((int64_t)blocks * kEwThreads) % (C / 8) == 0)
    mmr_launch((bn_bwd_apply_kernel<true>), blocks, kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(g), reinterpret_cast<const __nv_bfloat16*>(z), mean, invstd,
        coef, P, C, reinterpret_cast<__nv_bfloat16*>(dz));
  else
    mmr_launch((bn_bwd_apply_kernel<false>), blocks, kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(g), reinterpret_cast<const __nv_bfloat16*>(z), mean, invstd,
        coef, P, C, reinterpret_cast<__nv_bfloat16*>(dz));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_bn_bwd_apply_masked(const void* dx, const void* z, const float* mean, const float* invstd,
                                       const float* coef, const float* mask_scale, const float* mask_shift,
                                       int64_t P, int C, void* dz, mmr_stream_t stream) {
  MMR_REQUIRE(C % 8 == 0 && mask_scale && mask_shift, "C must be a multiple of 8; mask_scale / mask_shift required");
  if (ew16_ok(C, 5, dx, z, dz)) {
    const int64_t total = P * (C / 16);
    mmr_launch((bn_bwd_apply16_kernel<true>), ew_blocks(total, 32), kEwThreads,
               (size_t)(C / 16) * coef16_stride<5>() * sizeof(float), as_stream(stream),
               reinterpret_cast<const __nv_bfloat16*>(dx), reinterpret_cast<const __nv_bfloat16*>(z), mean, invstd, coef,
               mask_scale, mask_shift, total, C, reinterpret_cast<__nv_bfloat16*>(dz));
    MMR_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  const int blocks = ew_blocks((P * (C / 8) + 1) / 2, 16);
  if (((int64_t)blocks * kEwThreads) % (C / 8) == 0)
    mmr_launch((bn_bwd_apply_masked_kernel<true>), blocks, kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(dx), reinterpret_cast<const __nv_bfloat16*>(z), mean, invstd, coef,
        mask_scale, mask_shift, P, C, reinterpret_cast<__nv_bfloat16*>(dz));
  else
    mmr_launch((bn_bwd_apply_masked_kernel<false>), blocks, kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(dx), reinterpret_cast<const __nv_bfloat16*>(z), mean, invstd, coef,
        mask_scale, mask_shift, P, C, reinterpret_cast<__nv_bfloat16*>(dz));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_grad_gather(const MmrContrib* contribs, int ncontrib, const void* act, int N,
                               int H, int W, int C, void* g, double* partial, int nblk,
                               mmr_stream_t stream) {
  if (check_rows_layout(C)) return -1;
  ContribList cl;
  if (fill_contribs(cl, contribs, ncontrib)) return -1;
  const int64_t P = (int64_t)N * H * W;
  MMR_REQUIRE(P < ((int64_t)1 << 31), "row count must be below 2^31");
  launch_reduce<2>(nblk, as_stream(stream), nullptr, (uint32_t)P, C, partial, cl,
                   reinterpret_cast<const __nv_bfloat16*>(act), nullptr, nullptr, make_geom(H, W),
                   reinterpret_cast<__nv_bfloat16*>(g));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// Bias gradient of a conv without BatchNorm: column sums left by mmr_grad_gather.
__global__ void bias_grad_finalize_kernel(const double* __restrict__ partial, int nblk, int C,
                                          float* dbias, int accumulate) {
  pdl_prologue();
  double s1, s2;
  reduce_partials8(partial, nblk, C, blockIdx.x * 8, s1, s2);
  const int c = blockIdx.x * 8 + (threadIdx.x & 7);
  if ((threadIdx.x >> 3) != 0 || c >= C) return;
  dbias[c] = (accumulate ? dbias[c] : 0.f) + (float)s1;
}

extern "C" int mmr_bias_grad_finalize(const double* partial, int nblk, int C, float* dbias,
                                      int accumulate, mmr_stream_t stream) {
  mmr_launch((bias_grad_finalize_kernel), (C + 7) / 8, 256, 0, as_stream(stream), partial, nblk, C, dbias,
                                                                         accumulate);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_maxpool3x3s2_fwd(const void* x, int N, int H, int W, int C, void* out,
                                    uint8_t* idx, mmr_stream_t stream) {
  MMR_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int64_t total = (int64_t)N * Ho * Wo * (C / 8);
  MMR_REQUIRE(total < ((int64_t)1 << 31), "max-pool: tensor too large for 32-bit indexing");
  mmr_launch((maxpool_fwd_kernel), ew_blocks(total, 64), kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(x), N, H, W, C, reinterpret_cast<__nv_bfloat16*>(out),
      idx);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_maxpool3x3s2_bwd(const MmrContrib* contribs, int ncontrib, const uint8_t* idx,
                                    int N, int H, int W, int C, void* gin, mmr_stream_t stream) {
  MMR_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
  ContribList cl;
  if (fill_contribs(cl, contribs, ncontrib)) return -1;
  const int64_t total = (int64_t)N * H * W * (C / 8);
  MMR_REQUIRE(total < ((int64_t)1 << 31), "maxpool backward: tensor too large for 32-bit indexing");
  bool pooled = false;
  for (int i = 0; i < cl.n; ++i) pooled |= cl.pool2[i] != 0;
  const int blocks = ew_blocks(total, 64);
  __nv_bfloat16* go = reinterpret_cast<__nv_bfloat16*>(gin);
  if (!pooled && H % 2 == 0 && W % 2 == 0 && cl.n == 1)
    mmr_launch((maxpool_bwd_block_kernel<1>), ew_blocks(total / 4, 64), kEwThreads, 0, as_stream(stream), cl, idx, N, H, W, C, go);
  else if (!pooled && H % 2 == 0 && W % 2 == 0 && cl.n == 2)
    mmr_launch((maxpool_bwd_block_kernel<2>), ew_blocks(total / 4, 64), kEwThreads, 0, as_stream(stream), cl, idx, N, H, W, C, go);
  else if (!pooled && cl.n == 1)
    mmr_launch((maxpool_bwd_kernel<1>), blocks, kEwThreads, 0, as_stream(stream), cl, idx, N, H, W, C, go);
  else if (!pooled && cl.n == 2)
    mmr_launch((maxpool_bwd_kernel<2>), blocks, kEwThreads, 0, as_stream(stream), cl, idx, N, H, W, C, go);
  else if (!pooled && cl.n == 3)
    mmr_launch((maxpool_bwd_kernel<3>), blocks, kEwThreads, 0, as_stream(stream), cl, idx, N, H, W, C, go);
  else
    mmr_launch((maxpool_bwd_kernel<0>), blocks, kEwThreads, 0, as_stream(stream), cl, idx, N, H, W, C, go);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_maxpool2x2s2_fwd(const void* x, int N, int H, int W, int C, void* out, uint8_t* idx,
                                    mmr_stream_t stream) {
  MMR_REQUIRE(C % 8 == 0 && H >= 2 && W >= 2, "C must be a multiple of 8 and the image at least 2x2");
  const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
  MMR_REQUIRE((int64_t)N * H * W * (C / 8) < ((int64_t)1 << 31), "max-pool: tensor too large for 32-bit indexing");
  mmr_launch((maxpool2_fwd_kernel), ew_blocks(total, 64), kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(x), N, H, W, C, reinterpret_cast<__nv_bfloat16*>(out), idx);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_maxpool2x2s2_bwd(const MmrContrib* contribs, int ncontrib, const uint8_t* idx, int N, int H,
                                    int W, int C, void* gin, mmr_stream_t stream) {
  MMR_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
  ContribList cl;
  if (fill_contribs(cl, contribs, ncontrib)) return -1;
  const int64_t total = (int64_t)N * H * W * (C / 8);
  MMR_REQUIRE(total < ((int64_t)1 << 31), "max-pool backward: tensor too large for 32-bit indexing");
  mmr_launch((maxpool2_bwd_kernel), ew_blocks(total, 64), kEwThreads, 0, as_stream(stream), cl, idx, N, H, W, C, reinterpret_cast<__nv_bfloat16*>(gin));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_head_grad_prep(const float* dlogits, int N, int C, int H, int W, void* out,
                                  int cpad, float* dbias, int accumulate, mmr_stream_t stream) {
  MMR_REQUIRE(cpad % 8 == 0 && cpad >= C && C <= 16, "head_grad_prep: need C <= 16 <= cpad (mult of 8)");
  if (dbias && g_head_ws == nullptr) {
    MMR_CUDA_CHECK(cudaMalloc(&g_head_ws, sizeof(double) * 16 * kHeadBlocks));
  }
  const int64_t hw = (int64_t)H * W;
  if (cpad == 16 && hw % 4 == 0 && ((reinterpret_cast<uintptr_t>(dlogits) & 15) | (reinterpret_cast<uintptr_t>(out) & 31)) == 0) {
    if (C <= 4)
      mmr_launch((head_grad_prep4_kernel<4>), kHeadBlocks, kEwThreads, 0, as_stream(stream), dlogits, N, C, hw,
                 reinterpret_cast<__nv_bfloat16*>(out), dbias ? g_head_ws : nullptr);
    else
      mmr_launch((head_grad_prep4_kernel<16>), kHeadBlocks, kEwThreads, 0, as_stream(stream), dlogits, N, C, hw,
                 reinterpret_cast<__nv_bfloat16*>(out), dbias ? g_head_ws : nullptr);
  } else {
    mmr_launch((head_grad_prep_kernel), kHeadBlocks, kEwThreads, 0, as_stream(stream), dlogits, N, C, H, W, reinterpret_cast<__nv_bfloat16*>(out), cpad, dbias ? g_head_ws : nullptr);
  }
  MMR_CUDA_CHECK(cudaGetLastError());
  if (dbias) {
    mmr_launch((head_bias_finalize_kernel), 1, 256, 0, as_stream(stream), g_head_ws, kHeadBlocks, C, dbias,
                                                               accumulate);
    MMR_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

// Strip length of the row-walking bilinear kernels: as long as possible (the rows a strip shares with its
// neighbours are read twice) while the launch still fills the SMs several times over.
static int bilinear_strip_rows(int64_t threads_per_row_set, int H, int longest) {
  int rows = longest;
  while (rows > 2 && threads_per_row_set * ((H + rows - 1) / rows) < (int64_t)num_sms() * 2048 * 2) rows /= 2;
  return rows;
}

extern "C" int mmr_upsample_bilinear2x_fwd(const void* x, int N, int H, int W, int C, void* out,
                                           mmr_stream_t stream) {
  MMR_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
  MMR_REQUIRE(N >= 1 && N <= 65535 && H >= 1 && W >= 1, "bad shape %d x %d x %d", N, H, W);
  const int64_t pieces = (int64_t)2 * W * (C / 8);
  const int rows = bilinear_strip_rows(pieces * N, 2 * H, 32);
  dim3 grid((unsigned)((pieces + kEwThreads - 1) / kEwThreads), (unsigned)((2 * H + rows - 1) / rows), (unsigned)N);
  mmr_launch((upsample_bilinear2x_fwd_kernel), grid, kEwThreads, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(x), H, W, C, rows, reinterpret_cast<__nv_bfloat16*>(out));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_upsample_bilinear2x_bwd(const MmrContrib* contribs, int ncontrib, int N, int H, int W,
                                           int C, void* gin, mmr_stream_t stream) {
  MMR_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
  MMR_REQUIRE(N >= 1 && N <= 65535 && H >= 1 && W >= 1, "bad shape %d x %d x %d", N, H, W);
  ContribList cl;
  if (fill_contribs(cl, contribs, ncontrib)) return -1;
  const int64_t pieces = (int64_t)W * (C / 8);
  const int rows = bilinear_strip_rows(pieces * N, H, 16);
  dim3 grid((unsigned)((pieces + kEwThreads - 1) / kEwThreads), (unsigned)((H + rows - 1) / rows), (unsigned)N);
  if (cl.n == 1 && !cl.pool2[0])
    mmr_launch((upsample_bilinear2x_bwd_kernel<true>), grid, kEwThreads, 0, as_stream(stream), cl, H, W, C, rows, reinterpret_cast<__nv_bfloat16*>(gin));
  else
    mmr_launch((upsample_bilinear2x_bwd_kernel<false>), grid, kEwThreads, 0, as_stream(stream), cl, H, W, C, rows, reinterpret_cast<__nv_bfloat16*>(gin));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_upsample_nearest_f32_nchw(const float* in, int64_t planes, int h, int w, int f, float* out,
                                             mmr_stream_t stream) {
  MMR_REQUIRE(f >= 1 && f <= 32, "factor must be 1..32");
  const int64_t rows = planes * h * f, vec = rows * ((int64_t)w * f / 4);
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) & 7) | (reinterpret_cast<uintptr_t>(out) & 15)) == 0;
  if (aligned && vec < (int64_t)1 << 31 && f == 2 && w % 2 == 0)
    mmr_launch((upsample_nearest_f32_vec4_kernel<1>), ew_blocks(vec, 16), kEwThreads, 0, as_stream(stream), in, (uint32_t)rows, h, w, f, out);
  else if (aligned && vec < (int64_t)1 << 31 && f % 4 == 0)
    mmr_launch((upsample_nearest_f32_vec4_kernel<0>), ew_blocks(vec, 16), kEwThreads, 0, as_stream(stream), in, (uint32_t)rows, h, w, f, out);
  else
    mmr_launch((upsample_nearest_f32_kernel), ew_blocks(planes * h * w * f * f, 16), kEwThreads, 0, as_stream(stream), in, planes, h, w, f, out);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_sumpool_f32_nchw(const float* in, int64_t planes, int h, int w, int f, float* out,
                                    mmr_stream_t stream) {
  MMR_REQUIRE(f >= 1 && f <= 32, "factor must be 1..32");
  mmr_launch((sumpool_f32_kernel), ew_blocks(planes * h * w, 16), kEwThreads, 0, as_stream(stream), in, planes, h, w, f,
                                                                                         out);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_bn_bwd_reduce_fused(const MmrContrib* contribs, int ncontrib, const void* act, const void* z,
                                       const float* mean, const float* invstd, int N, int H, int W, int C,
                                       void* g, double* slots, int nblk, const float* gamma, float* dgamma,
                                       float* dbeta, int accumulate, float* coef, uint32_t* ticket,
                                       const float* mask_scale, const float* mask_shift, mmr_stream_t stream) {
  if (check_rows_layout(C)) return -1;
  MMR_REQUIRE(slots && coef && ticket, "fused backward reduce needs slots, coef and a ticket");
  MMR_REQUIRE((mask_scale == nullptr) == (mask_shift == nullptr) && !(mask_scale && act),
              "pass either the activation or (mask_scale, mask_shift)");
  ContribList cl;
  if (fill_contribs(cl, contribs, ncontrib)) return -1;
  const int64_t P = (int64_t)N * H * W;
  MMR_REQUIRE(P < ((int64_t)1 << 31), "row count must be below 2^31");
  BwdFused fz{gamma, dgamma, dbeta, coef, ticket, accumulate, mask_scale, mask_shift};
  if (mask_scale)
    launch_reduce<3>(nblk, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(z), (uint32_t)P, C, slots, cl,
                     nullptr, mean, invstd, make_geom(H, W), reinterpret_cast<__nv_bfloat16*>(g), fz);
  else
    launch_reduce<1>(nblk, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(z), (uint32_t)P, C, slots, cl,
                     reinterpret_cast<const __nv_bfloat16*>(act), mean, invstd, make_geom(H, W),
                     reinterpret_cast<__nv_bfloat16*>(g), fz);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}
