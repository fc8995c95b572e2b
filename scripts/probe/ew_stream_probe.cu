// Stand-alone probe: how should the HBM-bound BatchNorm-backward apply pass be shaped on B200?
//   dz = ca*g' + pz*z + q,  g' = (z*msc + msh > 0) ? dx : 0     (3 bf16 streams: 2 read, 1 written)
// Variants: bytes per load (16 / 32), independent items in flight per thread (U), per-channel constants in
// registers (thread keeps its channel group for the whole grid-stride loop) or in shared memory, CTAs per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ew_stream_probe ew_stream_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

template <int W> struct Vec;
template <> struct Vec<4> { uint32_t v[4]; };
template <> struct Vec<8> { uint32_t v[8]; };

__device__ __forceinline__ Vec<4> ldv(const Vec<4>* p) {
  Vec<4> r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "l"(p));
  return r;
}
__device__ __forceinline__ Vec<8> ldv(const Vec<8>* p) {
  Vec<8> r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                 "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ void stv(Vec<4>* p, const Vec<4>& r) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]) : "memory");
}
__device__ __forceinline__ void stv(Vec<8>* p, const Vec<8>& r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]),
               "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]) : "memory");
}
__device__ __forceinline__ float2 unpack(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// W = 32-bit words per load (4 -> 8 channels, 8 -> 16 channels); coef: [5][C] = ca, pz, q, msc, msh
template <int W, int U, bool SMEM>
__global__ void __launch_bounds__(256) apply_kernel(const Vec<W>* __restrict__ dx, const Vec<W>* __restrict__ z,
                                                    const float* __restrict__ coef, int64_t total, int C,
                                                    Vec<W>* __restrict__ dz) {
  extern __shared__ float sc[];
  constexpr int E = 2 * W;
  const int groups = C / E;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float k[SMEM ? 1 : 5][SMEM ? 1 : E];
  const int c0 = (int)(i % groups) * E;
  if (SMEM) {
    for (int t = threadIdx.x; t < 5 * C; t += blockDim.x) sc[t] = coef[t];
    __syncthreads();
  } else {
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int j = 0; j < E; ++j) k[SMEM ? 0 : a][SMEM ? 0 : j] = coef[a * C + c0 + j];
  }
  for (; i < total; i += U * stride) {
    Vec<W> g[U], zz[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < total) {
        g[u] = ldv(dx + i + u * stride);
        zz[u] = ldv(z + i + u * stride);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < total) {
        Vec<W> o;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const float2 gv = unpack(g[u].v[w]), zv = unpack(zz[u].v[w]);
          float ca0, ca1, pz0, pz1, q0, q1, ms0, ms1, mh0, mh1;
          if (SMEM) {
            const float2 a = *reinterpret_cast<const float2*>(sc + c0 + 2 * w), b = *reinterpret_cast<const float2*>(sc + C + c0 + 2 * w),
                         c = *reinterpret_cast<const float2*>(sc + 2 * C + c0 + 2 * w), d = *reinterpret_cast<const float2*>(sc + 3 * C + c0 + 2 * w),
                         e = *reinterpret_cast<const float2*>(sc + 4 * C + c0 + 2 * w);
            ca0 = a.x, ca1 = a.y, pz0 = b.x, pz1 = b.y, q0 = c.x, q1 = c.y, ms0 = d.x, ms1 = d.y, mh0 = e.x, mh1 = e.y;
          } else {
            ca0 = k[0][SMEM ? 0 : 2 * w], ca1 = k[0][SMEM ? 0 : 2 * w + 1];
            pz0 = k[SMEM ? 0 : 1][SMEM ? 0 : 2 * w], pz1 = k[SMEM ? 0 : 1][SMEM ? 0 : 2 * w + 1];
            q0 = k[SMEM ? 0 : 2][SMEM ? 0 : 2 * w], q1 = k[SMEM ? 0 : 2][SMEM ? 0 : 2 * w + 1];
            ms0 = k[SMEM ? 0 : 3][SMEM ? 0 : 2 * w], ms1 = k[SMEM ? 0 : 3][SMEM ? 0 : 2 * w + 1];
            mh0 = k[SMEM ? 0 : 4][SMEM ? 0 : 2 * w], mh1 = k[SMEM ? 0 : 4][SMEM ? 0 : 2 * w + 1];
          }
          const float g0 = fmaf(zv.x, ms0, mh0) > 0.f ? gv.x : 0.f, g1 = fmaf(zv.y, ms1, mh1) > 0.f ? gv.y : 0.f;
          o.v[w] = pack(ca0 * g0 + pz0 * zv.x + q0, ca1 * g1 + pz1 * zv.y + q1);
        }
        stv(dz + i + u * stride, o);
      }
  }
}

template <int W, int U, bool SMEM>
void run(const char* name, void* dx, void* z, float* coef, int64_t P, int C, void* dz, int per_sm, int sms) {
  const int64_t total = P * C / (2 * W);
  int64_t blocks = (total + 256 * U - 1) / (256 * U);
  if (per_sm > 0 && blocks > (int64_t)per_sm * sms) blocks = (int64_t)per_sm * sms;
  const int groups = C / (2 * W);
  if ((blocks * 256) % groups) { printf("%-28s skipped\n", name); return; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t smem = SMEM ? 5 * C * sizeof(float) : 0;
  for (int it = 0; it < 3; ++it)
    apply_kernel<W, U, SMEM><<<(int)blocks, 256, smem>>>((const Vec<W>*)dx, (const Vec<W>*)z, coef, total, C, (Vec<W>*)dz);
  cudaEventRecord(e0);
  const int iters = 20;
  for (int it = 0; it < iters; ++it)
    apply_kernel<W, U, SMEM><<<(int)blocks, 256, smem>>>((const Vec<W>*)dx, (const Vec<W>*)z, coef, total, C, (Vec<W>*)dz);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, apply_kernel<W, U, SMEM>, 256, smem);
  cudaError_t err = cudaGetLastError();
  printf("%-28s C %3d per_sm %3d blocks %7lld occ %d  %7.1f us  %5.2f TB/s %s\n", name, C, per_sm, (long long)blocks, nb,
         ms * 1e3, 3.0 * P * C * 2 / ms / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int64_t bytes = 16ll * 512 * 512 * 16 * 2;   // 134 MB per stream
  void *dx, *z, *dz;
  float* coef;
  cudaMalloc(&dx, bytes);
  cudaMalloc(&z, bytes);
  cudaMalloc(&dz, bytes);
  cudaMalloc(&coef, 5 * 512 * sizeof(float));
  cudaMemset(dx, 0x3c, bytes);
  cudaMemset(z, 0x3c, bytes);
  cudaMemset(coef, 0, 5 * 512 * sizeof(float));
  // plain device-to-device copy of one stream as the yardstick (read + write)
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) cudaMemcpyAsync(dz, dx, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int it = 0; it < 20; ++it) cudaMemcpyAsync(dz, dx, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D 134 MB: %.1f us  %.2f TB/s (read + write)\n", ms / 20 * 1e3, 2.0 * bytes / (ms / 20) / 1e9);
  }
  for (int C : {64, 16}) {
    const int64_t P = bytes / 2 / C;
    for (int per_sm : {8, 16, 32, 0}) {
      run<4, 1, false>("16B U1 regs", dx, z, coef, P, C, dz, per_sm, sms);
      run<4, 2, false>("16B U2 regs (current)", dx, z, coef, P, C, dz, per_sm, sms);
      run<4, 4, false>("16B U4 regs", dx, z, coef, P, C, dz, per_sm, sms);
      run<4, 2, true>("16B U2 smem", dx, z, coef, P, C, dz, per_sm, sms);
      run<4, 4, true>("16B U4 smem", dx, z, coef, P, C, dz, per_sm, sms);
      run<4, 8, true>("16B U8 smem", dx, z, coef, P, C, dz, per_sm, sms);
      run<8, 1, true>("32B U1 smem", dx, z, coef, P, C, dz, per_sm, sms);
      run<8, 2, true>("32B U2 smem", dx, z, coef, P, C, dz, per_sm, sms);
      run<8, 4, true>("32B U4 smem", dx, z, coef, P, C, dz, per_sm, sms);
    }
  }
  return 0;
}
