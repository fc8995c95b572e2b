// Optimiser-side kernels: Adam / AdamW over the flat fp32 parameter buffer and the gradient
// norm.
// All HBM-bound (Adam: 28 B/param).
#include "common.h"

namespace mmr {

constexpr int kOptThreads = 256;

// torch.optim.Adam (SU/ModelTraining.py:366: lr, weight_decay -> L2 added to the gradient)
// and AdamW (ED/Main_MMR_SegModel.py:878-880: decoupled decay), single-tensor formulas:
//   g = g*grad_scale (+ wd*p for Adam);  p *= (1 - lr*wd) for AdamW
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g
//   p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(kOptThreads)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
            float bc1, float bc2, int mode, float grad_scale) {
  pdl_prologue();
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = reinterpret_cast<float*>(&pp);
    const float* ga = reinterpret_cast<const float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gr = ga[j] * grad_scale;
      if (mode == 0)
        gr += wd * pa[j];
      else
        pa[j] *= (1.f - lr * wd);
      ma[j] = b1 * ma[j] + (1.f - b1) * gr;
      va[j] = b2 * va[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(va[j]) * inv_sqrt_bc2 + eps;
      pa[j] -= step_size * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float gr = g[i] * grad_scale;
    float pv = p[i];
    if (mode == 0)
      gr += wd * pv;
    else
      pv *= (1.f - lr * wd);
    const float mv = b1 * m[i] + (1.f - b1) * gr;
    const float vv = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mv;
    v[i] = vv;
    p[i] = pv - step_size * (mv / (sqrtf(vv) * inv_sqrt_bc2 + eps));
  }
}

// torch.optim.SGD(momentum, weight_decay) (SU/ModelTraining.py:372, 381), dampening 0, no Nesterov:
//   g = g*grad_scale + wd*p;  buf = first ? g : momentum*buf + g;  p -= lr*buf   (20 B/param)
__global__ void __launch_bounds__(kOptThreads)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, int64_t n, float lr,
           float momentum, float wd, int first, float grad_scale) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gr = g[i] * grad_scale + wd * pv;
    float b = gr;
    if (momentum != 0.f) {
      b = first ? gr : momentum * buf[i] + gr;
      buf[i] = b;
    }
    p[i] = pv - lr * b;
  }
}

__global__ void __launch_bounds__(kOptThreads)
sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  pdl_prologue();
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(g + i);
    acc += (double)x * (double)x;
  }
  __shared__ double sh[kOptThreads];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kOptThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, sh[0]);
}

// clip_grad_norm_'s scaling half: every thread reads the device-side sum of squares.
__global__ void __launch_bounds__(kOptThreads)
clip_scale_kernel(float* __restrict__ g, int64_t n, const double* __restrict__ sumsq, float max_norm,
                  float* __restrict__ norm_out) {
  pdl_prologue();
  const float total = (float)sqrt(*sumsq);
  if (norm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
  const float coef = max_norm / (total + 1e-6f);
  if (!(coef < 1.f)) return;
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(g)[i];
    v.x *= coef; v.y *= coef; v.z *= coef; v.w *= coef;
    reinterpret_cast<float4*>(g)[i] = v;
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    g[i] *= coef;
}

// Eval-mode BatchNorm folding, one CTA per job (a job is one BatchNorm layer: at most 2048 channels).
__global__ void __launch_bounds__(kOptThreads)
bn_fold_kernel(const MmrBnFoldJob* __restrict__ jobs, float eps) {
  pdl_prologue();
  const MmrBnFoldJob j = jobs[blockIdx.x];
  for (int c = threadIdx.x; c < j.C; c += blockDim.x) {
    const float sc = j.gamma[c] * rsqrtf(j.running_var[c] + eps);
    float mean = j.running_mean[c];
    if (j.conv_bias != nullptr) mean -= j.conv_bias[c];
    const float sh = j.beta[c] - mean * sc;
    for (int r = 0; r < j.rep; ++r) {
      j.scale[r * j.C + c] = sc;
      j.shift[r * j.C + c] = sh;
    }
  }
}

}  // namespace mmr

using namespace mmr;

extern "C" int mmr_clip_scale(float* g, int64_t n, const double* sumsq, float max_norm, float* norm_out,
                              mmr_stream_t stream) {
  MMR_REQUIRE(g && sumsq && n >= 0, "bad argument");
  MMR_REQUIRE(reinterpret_cast<uintptr_t>(g) % 16 == 0, "gradient buffer must be 16-byte aligned");
  int64_t blocks = (n / 4 + kOptThreads - 1) / kOptThreads;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  mmr_launch((clip_scale_kernel), (int)blocks, kOptThreads, 0, as_stream(stream), g, n, sumsq, max_norm, norm_out);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_bn_fold_batch(const MmrBnFoldJob* jobs_dev, int njobs, float eps, mmr_stream_t stream) {
  MMR_REQUIRE(jobs_dev && njobs >= 0, "bad argument");
  if (njobs == 0) return 0;
  mmr_launch((bn_fold_kernel), njobs, kOptThreads, 0, as_stream(stream), jobs_dev, eps);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                             float b1, float b2, float eps, float wd, float bc1, float bc2, int mode,
                             float grad_scale, mmr_stream_t stream) {
  MMR_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
               reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
              "adam buffers must be 16-byte aligned");
  int64_t blocks = (n / 4 + kOptThreads - 1) / kOptThreads;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  mmr_launch((adam_kernel), (int)blocks, kOptThreads, 0, as_stream(stream), p, g, m, v, n, lr, b1, b2, eps, wd,
                                                                  bc1, bc2, mode, grad_scale);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_zero_async(void* ptr, int64_t nbytes, mmr_stream_t stream) {
  MMR_REQUIRE(ptr != nullptr && nbytes >= 0, "bad argument");
  MMR_CUDA_CHECK(cudaMemsetAsync(ptr, 0, (size_t)nbytes, as_stream(stream)));
  return 0;
}

extern "C" int mmr_sgd_step(float* p, const float* g, float* momentum_buf, int64_t n, float lr, float momentum,
                            float wd, int first_step, float grad_scale, mmr_stream_t stream) {
  MMR_REQUIRE(p && g && n >= 0 && (momentum == 0.f || momentum_buf), "bad argument");
  if (n == 0) return 0;
  int64_t blocks = (n + kOptThreads - 1) / kOptThreads;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  mmr_launch((sgd_kernel), (int)blocks, kOptThreads, 0, as_stream(stream), p, g, momentum_buf, n, lr, momentum, wd,
             first_step, grad_scale);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_sumsq(const float* g, int64_t n, double* out, mmr_stream_t stream) {
  int64_t blocks = (n + kOptThreads * 4 - 1) / (kOptThreads * 4);
  const int64_t cap = (int64_t)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  mmr_launch((sumsq_kernel), (int)blocks, kOptThreads, 0, as_stream(stream), g, n, out);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}
