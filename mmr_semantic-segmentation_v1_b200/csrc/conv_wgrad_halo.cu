// Weight gradient of a 3x3 / stride 1 / pad 1 convolution on tcgen05, second kernel generation.
//
//   dW[co][ci][ky][kx] = sum over pixels  x[pixel + (ky-1, kx-1)][ci] * dz[pixel][co]
//
// As in conv_wgrad.cu both operands are pixel-major in memory (NHWC), i.e. MN-major UMMA operands,
// and the reduction runs over pixels.  What is new: the nine shifted activation boxes of a pixel
// tile are nine views of ONE halo tile (18 rows x (8*TX+2) pixels x one channel chunk) fetched by a
// single TMA box — the first generation issued one small TMA per tap and was bound by the rate at
// which one thread can issue them.  An M tile of 128 accumulator rows stacks 128/cb views:
//   cb = 64:  two taps per M tile (second atom = first + LBO), 5 M tiles, the last one half used;
//   cb = 32 / 16:  one M tile per filter row ky, atoms at 1-pixel steps (kx = 0,1,2 used).
// A CTA owns one (channel chunk, 64-wide output-channel slice) and a range of pixel macro tiles
// (split-K); fp32 partials are then summed in split order (deterministic) and scattered to OIHW.
// Nearest-x2 sources use the zero-stride replicating tensor maps of conv_halo.cu.
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace mmr {

constexpr int kWhThreads = 256;
constexpr int kWhMaxChunks = 16;
constexpr int kWhMaxStages = 6;
constexpr int kWhMaxAcc = 5;

#ifndef MMR_PREFETCH_DESC
#define MMR_PREFETCH_DESC 1
#endif
struct WhChunk {
  int32_t map, map_edge, c0, up, ci0;
  const uint8_t* base;  // mode 2 (cp.async gather): channel c0 of pixel (0, 0, 0) of the source
  int32_t pxb, sw, sh;  //   bytes per source pixel, stored width / height (half the conv's when up)
};

struct WhParams {
  const CUtensorMap* maps;  // device: [activation maps ...][dz map]
  WhChunk chunk[kWhMaxChunks];
  int nchunks, dzmap, dz_phased;  // dz_phased: N tile nt reads output phase nt through its own strided map
  int H, W, N, TX, tiles_x, tiles_y, total_tiles;
  int cb, xrb, bn, zrb, n_ntiles, A;
  int n_split, stages;
  int pitch[2];
  uint32_t x_tx_bytes[2], x_stage_bytes, z_tx_bytes, stage_bytes, tmem_cols;
  float* partial;  // [slice][split][prow][pcol]: mode 0 [A*128][bn], mode 1 [192][3*bn]
  float* dst;
  int dst_cout, dst_cin;
  int mode, prow, pcol;
  const uint8_t* dz_base;  // mode 2: dz pixel (0, 0, 0), channel 0
  int dz_pxb;
};

__device__ __forceinline__ bool wh_elect() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void wh_tma_load_5d(void* dst, const void* map, uint64_t* bar, int c0, int c1, int c2,
                                               int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void wh_umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\ntcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t wh_desc_hi(uint32_t sbo_bytes, uint32_t swz) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (swz << 29);
}

// View of accumulator a inside the halo tile: pixel offset of its first atom and the byte distance to
// the next atom along M (the LBO of the MN-major descriptor).
__device__ __forceinline__ void wh_view(int cb, int a, int pitch, int& voff_px, int& lbo_px) {
  if (cb == 64) {
    const int t0 = 2 * a, t1 = t0 + 1;
    const int o0 = (t0 / 3) * pitch + (t0 % 3);
    const int o1 = t1 < 9 ? (t1 / 3) * pitch + (t1 % 3) : o0 + 1;  // tap 9 does not exist: rows discarded
    voff_px = o0;
    lbo_px = o1 - o0;
  } else {
    voff_px = a * pitch;
    lbo_px = 1;
  }
}

__global__ void __launch_bounds__(kWhThreads, 1)
conv_wgrad_halo_kernel(const __grid_constant__ WhParams p) {
#if MMR_PREFETCH_DESC
  if (threadIdx.x == 96) {
    tma_prefetch_desc(&p.maps[p.dzmap + (p.dz_phased ? (int)(blockIdx.x % p.n_ntiles) : 0)]);
    tma_prefetch_desc(&p.maps[p.chunk[blockIdx.x / p.n_ntiles].map]);
    if (p.chunk[blockIdx.x / p.n_ntiles].up) tma_prefetch_desc(&p.maps[p.chunk[blockIdx.x / p.n_ntiles].map_edge]);
  }
#endif
  pdl_prologue_conv();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWhMaxStages;
  uint64_t* tmem_full = bars + 2 * kWhMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kWhMaxStages + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int slice = blockIdx.x;
  const int c = slice / p.n_ntiles, nt = slice % p.n_ntiles;
  const int split = blockIdx.y;
  const int k_begin = (int)(((long long)p.total_tiles * split) / p.n_split);
  const int k_end = (int)(((long long)p.total_tiles * (split + 1)) / p.n_split);

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  pdl_setup_done();
  const WhChunk ch = p.chunk[c];

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int tx = k_begin % p.tiles_x, t = k_begin / p.tiles_x;
    int ty = t % p.tiles_y, n = t / p.tiles_y;
    for (int it = k_begin; it < k_end; ++it) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (wh_elect()) {
        uint8_t* sx = smem + (size_t)stage * p.stage_bytes;
        uint8_t* sz = sx + p.x_stage_bytes;
        const int x0 = tx * 8 * p.TX, y0 = ty * 16;
        mbar_arrive_expect_tx(&full[stage], p.x_tx_bytes[ch.up] + p.z_tx_bytes);
        tma_load_4d(sz, &p.maps[p.dzmap + (p.dz_phased ? nt : 0)], &full[stage], p.dz_phased ? 0 : nt * p.bn, x0, y0, n);
        if (!ch.up) {
          tma_load_4d(sx, &p.maps[ch.map], &full[stage], ch.c0, x0 - 1, y0 - 1, n);
        } else {
          const int xl = (x0 >> 1) - 1, yl = y0 >> 1;
          const uint32_t rowb = (uint32_t)p.pitch[1] * p.xrb;
          wh_tma_load_5d(sx, &p.maps[ch.map_edge], &full[stage], ch.c0, 0, xl, yl - 1, n);
          wh_tma_load_5d(sx + rowb, &p.maps[ch.map], &full[stage], ch.c0, 0, xl, 0, n * (p.H >> 1) + yl);
          wh_tma_load_5d(sx + 17 * rowb, &p.maps[ch.map_edge], &full[stage], ch.c0, 0, xl, yl + 8, n);
        }
      }
      __syncwarp();
      if (++tx == p.tiles_x) {
        tx = 0;
        if (++ty == p.tiles_y) {
          ty = 0;
          ++n;
        }
      }
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, p.bn, 1, 1);  // both operands MN-major
    const uint32_t xrb = (uint32_t)p.xrb, zrb = (uint32_t)p.zrb;
    const uint32_t pitch = (uint32_t)p.pitch[ch.up];
    const uint32_t hiA = wh_desc_hi(pitch * xrb, swizzle_code(p.xrb));
    const uint32_t hiB = wh_desc_hi((uint32_t)(8 * p.TX) * zrb, swizzle_code(p.zrb));
    uint32_t a_const[kWhMaxAcc];
#pragma unroll
    for (int a = 0; a < kWhMaxAcc; ++a) {
      int voff, lbo;
      wh_view(p.cb, a, (int)pitch, voff, lbo);
      a_const[a] = (((uint32_t)voff * xrb) >> 4) + ((((uint32_t)lbo * xrb) >> 4) << 16);
    }
    const uint32_t smem0 = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = k_begin; it < k_end; ++it) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t x_base = smem0 + (uint32_t)stage * p.stage_bytes + (ch.up ? xrb : 0u);
      const uint32_t z_base = smem0 + (uint32_t)stage * p.stage_bytes + p.x_stage_bytes;
      if (wh_elect()) {
        for (int i = 0; i < p.TX; ++i) {
#pragma unroll 2
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t a_k = (x_base + ((uint32_t)(2 * kk) * pitch + 8u * i) * xrb) >> 4;
            const uint32_t b_k = ((z_base + ((uint32_t)(2 * kk * 8 * p.TX) + 8u * i) * zrb) >> 4) | 0x10000u;
            const uint32_t acc = (uint32_t)(it != k_begin || i != 0 || kk != 0);
#pragma unroll
            for (int a = 0; a < kWhMaxAcc; ++a)
              if (a < p.A) wh_umma(tmem_base + (uint32_t)(a * p.bn), a_k + a_const[a], hiA, b_k, hiB, idesc, acc);
          }
        }
      }
      __syncwarp();
      if (wh_elect()) umma_commit(&empty[stage]);
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (wh_elect()) {
      if (k_end > k_begin)
        umma_commit(tmem_full);
      else
        mbar_arrive(tmem_full);
    }
    __syncwarp();
    pdl_done();
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: fp32 partials
    const int q = warp - 4;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* base = p.partial + ((size_t)slice * p.n_split + split) * (size_t)(p.A * 128) * p.bn;
    for (int a = 0; a < p.A; ++a) {
      float* dst = base + ((size_t)a * 128 + q * 32 + lane) * p.bn;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * p.bn);
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        uint32_t r[16];
        if (k_end > k_begin) {
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}


// ---------------------------------------------------------------------------------------------------
// Third generation (mode 1, cb = 64): the filter COLUMN moves to the dz side.
//
//   dW[co][ci][ky][kx] = sum_q  x[q + (ky-1) rows][ci] * dz[q - (kx-1) pixels][co]
//
// so the B operand of one tcgen05.mma is three views of the dz tile one pixel apart (an MN-major operand
// with three atoms, LBO = one pixel: N = 3 * bn = 192) and the A operand stacks two filter ROWS of the
// 64-channel chunk (two atoms one tile row apart).  A K step of 16 pixels is 2 MMAs of N = 192 (tile 0 =
// rows ky 0 / 1, tile 1 = row ky 2 + 64 discarded lanes) instead of 5 MMAs of N = 64.  Why: an N = 64 MMA
// reads 4 KB of A + 2 KB of B from shared memory for 32 clk of tensor work and is bound by the 128 B / clk
// shared-memory port at 48 clk (scripts/probe/umma_rate_probe.cu: 48.0 clk at N = 64, 96.0 = the tensor
// floor at N = 192); per K step that is 30 KB / 240 clk before, 20 KB / 192 clk (tensor-bound) now.
// The activation tile needs no halo in x any more (18 rows x 8 TX pixels), the dz tile gets one
// (16 rows x (8 TX + 2) pixels); out-of-image pixels of either are zero-filled by TMA, which is exactly the
// convolution's zero padding.
template <int TX, int BN>
__global__ void __launch_bounds__(kWhThreads, 1)
conv_wgrad_kx_kernel(const __grid_constant__ WhParams p) {
  constexpr uint32_t PX = 8 * TX, PZ = 8 * TX + 2, XRB = 128, ZRB = BN * 2, NN = 3 * BN;
#if MMR_PREFETCH_DESC
  if (threadIdx.x == 96) {
    tma_prefetch_desc(&p.maps[p.dzmap + (p.dz_phased ? (int)(blockIdx.x % p.n_ntiles) : 0)]);
    tma_prefetch_desc(&p.maps[p.chunk[blockIdx.x / p.n_ntiles].map]);
    if (p.chunk[blockIdx.x / p.n_ntiles].up) tma_prefetch_desc(&p.maps[p.chunk[blockIdx.x / p.n_ntiles].map_edge]);
  }
#endif
  pdl_prologue_conv();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWhMaxStages;
  uint64_t* tmem_full = bars + 2 * kWhMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kWhMaxStages + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int slice = blockIdx.x;
  const int c = slice / p.n_ntiles, nt = slice % p.n_ntiles;
  const int split = blockIdx.y;
  const int k_begin = (int)(((long long)p.total_tiles * split) / p.n_split);
  const int k_end = (int)(((long long)p.total_tiles * (split + 1)) / p.n_split);

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  pdl_setup_done();
  const WhChunk ch = p.chunk[c];

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int tx = k_begin % p.tiles_x, t = k_begin / p.tiles_x;
    int ty = t % p.tiles_y, n = t / p.tiles_y;
    for (int it = k_begin; it < k_end; ++it) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (wh_elect()) {
        uint8_t* sx = smem + (size_t)stage * p.stage_bytes;
        uint8_t* sz = sx + p.x_stage_bytes;
        const int x0 = tx * (int)PX, y0 = ty * 16;
        mbar_arrive_expect_tx(&full[stage], p.x_tx_bytes[0] + p.z_tx_bytes);
        tma_load_4d(sz, &p.maps[p.dzmap], &full[stage], nt * BN, x0 - 1, y0, n);
        if (!ch.up) {
          tma_load_4d(sx, &p.maps[ch.map], &full[stage], ch.c0, x0, y0 - 1, n);
        } else {
          const int xl = x0 >> 1, yl = y0 >> 1;
          constexpr uint32_t rowb = PX * XRB;
          wh_tma_load_5d(sx, &p.maps[ch.map_edge], &full[stage], ch.c0, 0, xl, yl - 1, n);
          wh_tma_load_5d(sx + rowb, &p.maps[ch.map], &full[stage], ch.c0, 0, xl, 0, n * (p.H >> 1) + yl);
          wh_tma_load_5d(sx + 17 * rowb, &p.maps[ch.map_edge], &full[stage], ch.c0, 0, xl, yl + 8, n);
        }
      }
      __syncwarp();
      if (++tx == p.tiles_x) {
        tx = 0;
        if (++ty == p.tiles_y) {
          ty = 0;
          ++n;
        }
      }
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, NN, 1, 1);  // both operands MN-major
    const uint32_t hiA = wh_desc_hi(PX * XRB, 2u);
    const uint32_t hiB = wh_desc_hi(PZ * ZRB, ZRB == 128 ? 2u : 4u);
    constexpr uint32_t a_lbo = (PX * XRB) >> 4;  // next atom along M: the same pixels one tile row down (ky + 1)
    constexpr uint32_t b_lbo = ZRB >> 4;         // next atom along N: one pixel to the right (kx - 1)
    const uint32_t smem0 = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = k_begin; it < k_end; ++it) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t x_lo = ((smem0 + (uint32_t)stage * p.stage_bytes) >> 4) | (a_lbo << 16);
      const uint32_t z_lo = ((smem0 + (uint32_t)stage * p.stage_bytes + p.x_stage_bytes) >> 4) | (b_lbo << 16);
      if (wh_elect()) {
#pragma unroll
        for (int i = 0; i < TX; ++i) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t a_k = x_lo + (((uint32_t)(2 * kk) * PX + 8u * i) * XRB >> 4);
            const uint32_t b_k = z_lo + (((uint32_t)(2 * kk) * PZ + 8u * i) * ZRB >> 4);
            if (i == 0 && kk == 0) {
              const uint32_t acc = (uint32_t)(it != k_begin);
              wh_umma(tmem_base, a_k, hiA, b_k, hiB, idesc, acc);
              wh_umma(tmem_base + NN, a_k + (2u * PX * XRB >> 4), hiA, b_k, hiB, idesc, acc);
            } else {
              wh_umma(tmem_base, a_k, hiA, b_k, hiB, idesc, 1u);
              wh_umma(tmem_base + NN, a_k + (2u * PX * XRB >> 4), hiA, b_k, hiB, idesc, 1u);
            }
          }
        }
      }
      __syncwarp();
      if (wh_elect()) umma_commit(&empty[stage]);
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (wh_elect()) {
      if (k_end > k_begin)
        umma_commit(tmem_full);
      else
        mbar_arrive(tmem_full);
    }
    __syncwarp();
    pdl_done();
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: fp32 partials [192][NN]
    const int q = warp - 4;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* base = p.partial + ((size_t)slice * p.n_split + split) * (size_t)(192 * NN);
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
      if (t == 1 && q >= 2) break;  // tile 1: lanes 64..127 belong to no filter row
      float* dst = base + ((size_t)t * 128 + q * 32 + lane) * NN;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * NN;
#pragma unroll 1
      for (int c0 = 0; c0 < (int)NN; c0 += 16) {
        uint32_t r[16];
        if (k_end > k_begin) {
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}


// ---------------------------------------------------------------------------------------------------
// Mode 2: the same kx-on-dz formulation for the narrow layers (16- / 32-channel chunks, 16 / 32 output channels:
// x_0_4.conv1 / conv2, the head, x_0_3.conv2).  One M tile stacks 128 / CB filter rows of the chunk (atoms one
// tile row apart; rows ky >= 3 are discarded), N = 3 * BN: ONE tcgen05.mma per 16 pixels instead of three of
// N = BN.  TMA moves 32- / 64-byte pixel rows one at a time (~4 clk per row: 137 of the 165 us of these launches),
// so the tiles are gathered with 16-byte cp.async by the six warps that have nothing else to do during the main
// loop (TMEM allocator, the idle warp, the four epilogue warps): scripts/probe/ldgsts_rate_probe.cu measures
// 24 B/clk/SM = the HBM roofline with four or more warps.  Zero fill through src-size 0 is the convolution's
// padding; the destination chunk index is XORed with address bits [7, 7 + log2(chunks per pixel)) -- SWIZZLE_32B /
// 64B as tcgen05.mma expects it; completion through cp.async.mbarrier.arrive, the MMA warp crosses to the async
// proxy after its wait.
__device__ __forceinline__ void wh_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
constexpr int kWhThinProducers = 6;   // warps 2 .. 7

template <int TX, int CB, int BN>
__global__ void __launch_bounds__(kWhThreads, 1)
conv_wgrad_thin_kernel(const __grid_constant__ WhParams p) {
  constexpr uint32_t PX = 8 * TX, PZ = 8 * TX + 2, XRB = CB * 2, ZRB = BN * 2, NN = 3 * BN;
  constexpr uint32_t CX = XRB / 16, CZ = ZRB / 16;            // 16-byte chunks per pixel
  constexpr uint32_t XPR = PX * CX, ZPR = PZ * CZ;            // chunks per tile row
  constexpr uint32_t XTOT = 18 * XPR, ZTOT = 16 * ZPR;        // chunks per tile
  constexpr uint32_t LIVE = 3 * CB;                           // accumulator rows that belong to a filter row
  pdl_prologue_conv();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWhMaxStages;
  uint64_t* tmem_full = bars + 2 * kWhMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kWhMaxStages + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int slice = blockIdx.x;
  const int c = slice / p.n_ntiles, nt = slice % p.n_ntiles;
  const int split = blockIdx.y;
  const int k_begin = (int)(((long long)p.total_tiles * split) / p.n_split);
  const int k_end = (int)(((long long)p.total_tiles * (split + 1)) / p.n_split);

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 32 * kWhThinProducers);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  pdl_setup_done();
  const WhChunk ch = p.chunk[c];

  if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, NN, 1, 1);  // both operands MN-major
    const uint32_t hiA = wh_desc_hi(PX * XRB, XRB == 64 ? 4u : 6u);
    const uint32_t hiB = wh_desc_hi(PZ * ZRB, ZRB == 64 ? 4u : 6u);
    constexpr uint32_t a_lbo = (PX * XRB) >> 4;  // next atom along M: the same pixels one tile row down (ky + 1)
    constexpr uint32_t b_lbo = ZRB >> 4;         // next atom along N: one pixel to the right (kx - 1)
    const uint32_t smem0 = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = k_begin; it < k_end; ++it) {
      mbar_wait(&full[stage], phase);
      fence_proxy_async_smem();   // cp.async wrote the tiles through the generic proxy
      tc_fence_after();
      const uint32_t x_lo = ((smem0 + (uint32_t)stage * p.stage_bytes) >> 4) | (a_lbo << 16);
      const uint32_t z_lo = ((smem0 + (uint32_t)stage * p.stage_bytes + p.x_stage_bytes) >> 4) | (b_lbo << 16);
      if (wh_elect()) {
#pragma unroll
        for (int i = 0; i < TX; ++i) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t a_k = x_lo + (((uint32_t)(2 * kk) * PX + 8u * i) * XRB >> 4);
            const uint32_t b_k = z_lo + (((uint32_t)(2 * kk) * PZ + 8u * i) * ZRB >> 4);
            if (i == 0 && kk == 0)
              wh_umma(tmem_base, a_k, hiA, b_k, hiB, idesc, (uint32_t)(it != k_begin));
            else
              wh_umma(tmem_base, a_k, hiA, b_k, hiB, idesc, 1u);
          }
        }
      }
      __syncwarp();
      if (wh_elect()) umma_commit(&empty[stage]);
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (wh_elect()) {
      if (k_end > k_begin)
        umma_commit(tmem_full);
      else
        mbar_arrive(tmem_full);
    }
    __syncwarp();
    pdl_done();
  } else if (warp >= 2) {
    // ---------------------------------------------------------------- gather: warps 2 .. 7
    const int pw = warp - 2;
    int stage = 0;
    uint32_t phase = 0;
    int tx = k_begin % p.tiles_x, t = k_begin / p.tiles_x;
    int ty = t % p.tiles_y, n = t / p.tiles_y;
    const uint8_t* zimg0 = p.dz_base + (size_t)nt * BN * 2;
    for (int it = k_begin; it < k_end; ++it) {
      mbar_wait(&empty[stage], phase ^ 1);
      const uint32_t sx = smem_u32(smem + (size_t)stage * p.stage_bytes);
      const uint32_t sz = sx + p.x_stage_bytes;
      const int x0 = tx * (int)PX, y0 = ty * 16;
      const uint8_t* ximg = ch.base + (size_t)n * ch.sh * ch.sw * ch.pxb;
      const uint8_t* zimg = zimg0 + (size_t)n * p.H * p.W * p.dz_pxb;
      // activation tile: 18 rows (y0 - 1 ...) x PX pixels, no halo in x
      for (uint32_t e = (uint32_t)pw * 32 + lane; e < XTOT; e += 32 * kWhThinProducers) {
        const uint32_t r = e / XPR, q = e % XPR, px = q / CX, j = q % CX;
        int y = y0 - 1 + (int)r, x = x0 + (int)px;
        const bool ok = y >= 0 && y < p.H && x < p.W;
        if (ch.up) y >>= 1, x >>= 1;
        const uint32_t off = (r * PX + px) * XRB;
        wh_cp_async16(sx + off + ((j ^ ((off >> 7) & (CX - 1))) << 4),
                      ok ? ximg + ((size_t)y * ch.sw + x) * ch.pxb + j * 16 : ximg, ok ? 16u : 0u);
      }
      // dz tile: 16 rows x PZ pixels (x0 - 1 ...)
      for (uint32_t e = (uint32_t)pw * 32 + lane; e < ZTOT; e += 32 * kWhThinProducers) {
        const uint32_t r = e / ZPR, q = e % ZPR, px = q / CZ, j = q % CZ;
        const int y = y0 + (int)r, x = x0 - 1 + (int)px;
        const bool ok = y < p.H && x >= 0 && x < p.W;
        const uint32_t off = (r * PZ + px) * ZRB;
        wh_cp_async16(sz + off + ((j ^ ((off >> 7) & (CZ - 1))) << 4),
                      ok ? zimg + ((size_t)y * p.W + x) * p.dz_pxb + j * 16 : zimg, ok ? 16u : 0u);
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[stage])) : "memory");
      if (++tx == p.tiles_x) {
        tx = 0;
        if (++ty == p.tiles_y) {
          ty = 0;
          ++n;
        }
      }
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (warp >= 4) {
      // ---------------------------------------------------------------- epilogue: fp32 partials [3 CB][NN]
      const int q = warp - 4;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int row = q * 32 + lane;
      if (q * 32 < (int)LIVE) {   // warp-uniform: tcgen05.ld is a warp-wide instruction
        float* dst = p.partial + ((size_t)slice * p.n_split + split) * (size_t)(LIVE * NN) + (size_t)row * NN;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < (int)NN; c0 += 16) {
          uint32_t r[16];
          if (k_end > k_begin) {
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = 0u;
          }
          if (row < (int)LIVE) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// Sum the split-K partials in split order and scatter into OIHW fp32 (dst[co][ci][tap]).  One
// thread owns four consecutive output channels of one accumulator row (float4 loads, four splits in
// flight); rows that belong to no filter tap are skipped before any load.
__global__ void __launch_bounds__(256)
wgrad_halo_reduce_kernel(const __grid_constant__ WhParams p, int accumulate) {
  pdl_prologue();
  // NL split lanes x (256 / NL) outputs (float4 each) per CTA, NL = 8 / 4 / 2 / 1 by the split count (with
  // two splits, eight lanes left six of eight warps idle: 45 us on the 512-channel layers): a lane sums the
  // splits k = lane, lane + NL, ... with four loads in flight, the lane sums are added in lane order
  // through shared memory (deterministic), so the serial chain per thread is n_split / NL long.
  __shared__ float4 part[256];
  const int nl = p.n_split >= 8 ? 8 : (p.n_split >= 4 ? 4 : (p.n_split >= 2 ? 2 : 1));
  const int per = 256 / nl;
  const int pc4 = p.pcol >> 2;
  const size_t per_slice = (size_t)p.prow * p.pcol;
  const size_t total4 = (size_t)p.prow * pc4 * p.nchunks * p.n_ntiles;  // modes 0 / 1: a multiple of 256
  const int sl = threadIdx.x / per, o = threadIdx.x % per;
  for (size_t base = (size_t)blockIdx.x * per; base < total4; base += (size_t)gridDim.x * per) {
    const size_t idx = base + o < total4 ? base + o : total4 - 1;   // mode 2: total4 need not be a multiple of `per`
    const bool in_range = base + o < total4;
    const int n4 = (int)(idx % pc4);
    size_t t = idx / pc4;
    const int row = (int)(t % p.prow);
    const int slice = (int)(t / p.prow);
    const int c = slice / p.n_ntiles, nt = slice % p.n_ntiles;
    int tap, chn, co;
    if (p.mode >= 1) {  // rows (ky, ci), columns (2 - kx, co)
      const int col = n4 * 4;
      tap = (row / p.cb) * 3 + 2 - col / p.bn;
      chn = row % p.cb;
      co = nt * p.bn + col % p.bn;
    } else {
      const int a = row >> 7, r = row & 127;
      if (p.cb == 64) {
        tap = 2 * a + (r >> 6);
        chn = r & 63;
      } else {
        const int kx = r / p.cb;
        tap = kx < 3 ? 3 * a + kx : 9;
        chn = r % p.cb;
      }
      co = nt * p.bn + n4 * 4;
    }
    const int ci = p.chunk[c].ci0 + chn;
    const bool live = in_range && tap < 9 && co < p.dst_cout && ci < p.dst_cin;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      const float4* src = reinterpret_cast<const float4*>(p.partial + (size_t)slice * p.n_split * per_slice +
                                                          (size_t)row * p.pcol) + n4;
      const size_t step = per_slice >> 2;
      int k = sl;
      for (; k + 3 * nl < p.n_split; k += 4 * nl) {
        const float4 v0 = __ldg(src + (size_t)k * step), v1 = __ldg(src + (size_t)(k + nl) * step);
        const float4 v2 = __ldg(src + (size_t)(k + 2 * nl) * step), v3 = __ldg(src + (size_t)(k + 3 * nl) * step);
        s.x += v0.x; s.y += v0.y; s.z += v0.z; s.w += v0.w;
        s.x += v1.x; s.y += v1.y; s.z += v1.z; s.w += v1.w;
        s.x += v2.x; s.y += v2.y; s.z += v2.z; s.w += v2.w;
        s.x += v3.x; s.y += v3.y; s.z += v3.z; s.w += v3.w;
      }
      for (; k < p.n_split; k += nl) {
        const float4 v = __ldg(src + (size_t)k * step);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    if (nl > 1) {
      part[threadIdx.x] = s;
      __syncthreads();
    }
    if (sl == 0 && live) {
      for (int j = 1; j < nl; ++j) {
        const float4 v = part[j * per + o];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (co + j >= p.dst_cout) break;
        float* d = p.dst + ((size_t)(co + j) * p.dst_cin + ci) * 9 + tap;
        *d = accumulate ? *d + sv[j] : sv[j];
      }
    }
    if (nl > 1) __syncthreads();
  }
}

typedef CUresult (*WhEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
void* get_encode_tiled();  // common.cu

static int wh_encode(CUtensorMap* out, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                     const cuuint32_t* box, int inner_bytes, const char* what) {
  WhEncodeTiledFn enc = reinterpret_cast<WhEncodeTiledFn>(get_encode_tiled());
  MMR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  MMR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: base must be 16-byte aligned", what);
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                    : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s, rank %d) -> CUresult %d", what, rank, (int)r);
  return 0;
}

struct WhPlan {
  WhParams prm;
  void* dev_blob = nullptr;
  size_t smem_bytes = 0;
};

}  // namespace mmr

using namespace mmr;

extern "C" int64_t mmr_wgrad_halo_partial_floats(int nchunks, int cb, int bn, int n_ntiles, int n_split) {
  const int A = cb == 64 ? 5 : 3;
  return (int64_t)nchunks * n_ntiles * n_split * A * 128 * bn;
}

extern "C" int64_t mmr_wgrad_kx_partial_floats(int nchunks, int bn, int n_ntiles, int n_split) {
  return (int64_t)nchunks * n_ntiles * n_split * 192 * 3 * bn;
}

template <int TX, int BN>
static cudaError_t wh_kx_attr() {
  return cudaFuncSetAttribute(conv_wgrad_kx_kernel<TX, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
}

extern "C" int64_t mmr_wgrad_thin_partial_floats(int nchunks, int cb, int bn, int n_ntiles, int n_split) {
  return (int64_t)nchunks * n_ntiles * n_split * 3 * cb * 3 * bn;
}

// mode 2 instantiations: (tx, cb, bn) in {2, 4} x {16, 32} x {16, 32}
template <int TX, int CB, int BN>
static cudaError_t wh_thin_do(const WhParams* p, dim3 grid, size_t smem, cudaStream_t st) {
  if (!p)
    return cudaFuncSetAttribute(conv_wgrad_thin_kernel<TX, CB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(227 * 1024));
  return mmr_launch((conv_wgrad_thin_kernel<TX, CB, BN>), grid, kWhThreads, smem, st, *p);
}
static cudaError_t wh_thin_dispatch(int tx, int cb, int bn, const WhParams* p, dim3 grid, size_t smem, cudaStream_t st) {
#define MMR_THIN(TX_, CB_, BN_) \
  if (tx == TX_ && cb == CB_ && bn == BN_) return wh_thin_do<TX_, CB_, BN_>(p, grid, smem, st);
  MMR_THIN(2, 16, 16) MMR_THIN(2, 16, 32) MMR_THIN(2, 32, 16) MMR_THIN(2, 32, 32)
  MMR_THIN(4, 16, 16) MMR_THIN(4, 16, 32) MMR_THIN(4, 32, 16) MMR_THIN(4, 32, 32)
#undef MMR_THIN
  return cudaErrorInvalidValue;
}

extern "C" int mmr_wgrad_halo_plan_create(const MmrWgradHaloDesc* d, void** out_plan) {
  MMR_REQUIRE(d && out_plan, "null argument");
  MMR_REQUIRE(d->nsrc >= 1 && d->nsrc <= 6, "nsrc must be 1..6, got %d", d->nsrc);
  MMR_REQUIRE(d->cb == 64 || d->cb == 32 || d->cb == 16, "cb must be 16/32/64, got %d", d->cb);
  MMR_REQUIRE(d->bn == 64 || d->bn == 32 || d->bn == 16, "bn must be 16/32/64, got %d", d->bn);
  const int dz_channels = d->dz.up == 2 ? 4 * d->dz.C : d->dz.C;   // up == 2: four output phases (see below)
  MMR_REQUIRE(d->cout_gemm % d->bn == 0 && dz_channels >= d->cout_gemm,
              "cout_gemm %d must be a multiple of bn %d and fit dz (%d)", d->cout_gemm, d->bn, dz_channels);
  MMR_REQUIRE(d->tx == 1 || d->tx == 2 || d->tx == 4, "tx must be 1, 2 or 4");
  MMR_REQUIRE((d->dz.up == 1 || d->dz.up == 2) && d->dz.H == d->H && d->dz.W == d->W && d->dz.N == d->N,
              "dz resolution mismatch");
  MMR_REQUIRE(d->partial && d->dst, "null output");
  MMR_REQUIRE(d->mode >= 0 && d->mode <= 2, "mode must be 0 (taps on the activation side), 1 (filter columns on dz) "
              "or 2 (the same for narrow layers, cp.async gather)");
  const bool kx = d->mode == 1;
  const bool thin = d->mode == 2;
  if (thin)
    MMR_REQUIRE((d->cb == 16 || d->cb == 32) && (d->bn == 16 || d->bn == 32) && (d->tx == 2 || d->tx == 4) && d->dz.up == 1,
                "mode 2 needs cb = 16 / 32, bn = 16 / 32, tx = 2 / 4 and a plain dz (got cb %d bn %d tx %d)", d->cb, d->bn, d->tx);
  if (kx)
    MMR_REQUIRE(d->cb == 64 && (d->bn == 64 || d->bn == 32) && (d->tx == 1 || d->tx == 2) && d->dz.up == 1,
                "mode 1 needs cb = 64, bn = 64 / 32, tx = 1 / 2 and a plain dz (got cb %d bn %d tx %d)", d->cb, d->bn, d->tx);
  WhPlan* pl = new WhPlan();
  WhParams& p = pl->prm;
  memset(&p, 0, sizeof(p));
  p.H = d->H;
  p.W = d->W;
  p.N = d->N;
  p.TX = d->tx;
  p.cb = d->cb;
  p.xrb = d->cb * 2;
  p.bn = d->bn;
  p.zrb = d->bn * 2;
  p.n_ntiles = d->cout_gemm / d->bn;
  p.A = d->cb == 64 ? 5 : 3;
  p.tiles_x = (d->W + 8 * d->tx - 1) / (8 * d->tx);
  p.tiles_y = (d->H + 15) / 16;
  p.total_tiles = p.tiles_x * p.tiles_y * d->N;
  p.pitch[0] = (kx || thin) ? 8 * d->tx : 8 * d->tx + 2;
  p.pitch[1] = (kx || thin) ? 8 * d->tx : 8 * d->tx + 4;
  p.mode = d->mode;
  p.prow = thin ? 3 * d->cb : (kx ? 192 : p.A * 128);
  p.pcol = (kx || thin) ? 3 * d->bn : d->bn;
  p.dz_base = reinterpret_cast<const uint8_t*>(d->dz.ptr);
  p.dz_pxb = d->dz.C * 2;

  std::vector<CUtensorMap> maps;
  int nchunks = 0, ci0 = 0;
  bool any_up = false;
  for (int si = 0; si < d->nsrc; ++si) {
    const MmrHaloSrc& s = d->src[si];
    MMR_REQUIRE(s.C % d->cb == 0, "source %d: %d channels are not a multiple of the chunk %d", si, s.C, d->cb);
    MMR_REQUIRE(s.N == d->N, "source %d: batch mismatch", si);
    int mi = -1, me = -1;
    if (thin) {
      MMR_REQUIRE(s.up == 1 ? (s.H == d->H && s.W == d->W) : (s.up == 2 && s.H * 2 == d->H && s.W * 2 == d->W),
                  "source %d: resolution mismatch", si);
      any_up |= s.up == 2;
    } else if (s.up == 1) {
      MMR_REQUIRE(s.H == d->H && s.W == d->W, "source %d: resolution mismatch", si);
      cuuint64_t dims[4] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.N};
      cuuint64_t str[3] = {(cuuint64_t)s.C * 2, (cuuint64_t)s.C * 2 * s.W, (cuuint64_t)s.C * 2 * s.W * s.H};
      cuuint32_t box[4] = {(cuuint32_t)d->cb, (cuuint32_t)p.pitch[0], 18, 1};
      maps.emplace_back();
      mi = (int)maps.size() - 1;
      if (wh_encode(&maps[mi], s.ptr, 4, dims, str, box, p.xrb, "wgrad halo source")) { delete pl; return -1; }
    } else {
      MMR_REQUIRE(s.up == 2 && s.H * 2 == d->H && s.W * 2 == d->W, "source %d: upsampled resolution mismatch", si);
      MMR_REQUIRE(d->H % 16 == 0, "nearest-x2 sources need H %% 16 == 0 (got %d)", d->H);
      any_up = true;
      const cuuint32_t bw = (cuuint32_t)(kx ? 4 * d->tx : 4 * d->tx + 2);
      {
        cuuint64_t dims[5] = {(cuuint64_t)s.C, 2, (cuuint64_t)s.W, 2, (cuuint64_t)s.N * s.H};
        cuuint64_t str[4] = {0, (cuuint64_t)s.C * 2, 0, (cuuint64_t)s.C * 2 * s.W};
        cuuint32_t box[5] = {(cuuint32_t)d->cb, 2, bw, 2, 8};
        maps.emplace_back();
        mi = (int)maps.size() - 1;
        if (wh_encode(&maps[mi], s.ptr, 5, dims, str, box, p.xrb, "wgrad upsampled body")) { delete pl; return -1; }
      }
      {
        cuuint64_t dims[5] = {(cuuint64_t)s.C, 2, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.N};
        cuuint64_t str[4] = {0, (cuuint64_t)s.C * 2, (cuuint64_t)s.C * 2 * s.W, (cuuint64_t)s.C * 2 * s.W * s.H};
        cuuint32_t box[5] = {(cuuint32_t)d->cb, 2, bw, 1, 1};
        maps.emplace_back();
        me = (int)maps.size() - 1;
        if (wh_encode(&maps[me], s.ptr, 5, dims, str, box, p.xrb, "wgrad upsampled edge")) { delete pl; return -1; }
      }
    }
    for (int c0 = 0; c0 < s.C; c0 += d->cb) {
      MMR_REQUIRE(nchunks < kWhMaxChunks, "more than %d channel chunks", kWhMaxChunks);
      p.chunk[nchunks++] = WhChunk{mi, me, c0, s.up == 2 ? 1 : 0, ci0 + c0,
                                   reinterpret_cast<const uint8_t*>(s.ptr) + (size_t)c0 * 2, s.C * 2, s.W, s.H};
    }
    ci0 += s.C;
  }
  p.nchunks = nchunks;
  {
    const MmrHaloSrc& z = d->dz;
    cuuint32_t box[4] = {(cuuint32_t)d->bn, (cuuint32_t)(kx ? 8 * d->tx + 2 : 8 * d->tx), 16, 1};
    p.dzmap = (int)maps.size();
    p.dz_phased = z.up == 2;
    if (thin) {
      // gathered with cp.async: no tensor map
    } else if (!p.dz_phased) {
      cuuint64_t dims[4] = {(cuuint64_t)z.C, (cuuint64_t)z.W, (cuuint64_t)z.H, (cuuint64_t)z.N};
      cuuint64_t str[3] = {(cuuint64_t)z.C * 2, (cuuint64_t)z.C * 2 * z.W, (cuuint64_t)z.C * 2 * z.W * z.H};
      maps.emplace_back();
      if (wh_encode(&maps[p.dzmap], z.ptr, 4, dims, str, box, p.zrb, "wgrad dz")) { delete pl; return -1; }
    } else {
      // dz is [N][2H][2W][C] and the GEMM's output channel (q, c) is channel c of pixel (2y + qy, 2x + qx): the
      // space-to-depth stem (mmr_stem_s2d_*).  One strided map per phase = per N tile.
      MMR_REQUIRE(d->bn == z.C && d->cout_gemm == 4 * z.C, "phased dz needs bn = C and cout_gemm = 4 C");
      const cuuint64_t pxb = (cuuint64_t)z.C * 2, rowb = pxb * z.W * 2;
      for (int q = 0; q < 4; ++q) {
        cuuint64_t dims[4] = {(cuuint64_t)z.C, (cuuint64_t)z.W, (cuuint64_t)z.H, (cuuint64_t)z.N};
        cuuint64_t str[3] = {pxb * 2, rowb * 2, rowb * 2 * z.H};
        const uint8_t* base = reinterpret_cast<const uint8_t*>(z.ptr) + (size_t)(q >> 1) * rowb + (size_t)(q & 1) * pxb;
        maps.emplace_back();
        if (wh_encode(&maps[p.dzmap + q], base, 4, dims, str, box, p.zrb, "wgrad dz phase")) { delete pl; return -1; }
      }
    }
  }
  p.x_tx_bytes[0] = (uint32_t)(18 * p.pitch[0] * p.xrb);
  p.x_tx_bytes[1] = (uint32_t)(18 * p.pitch[1] * p.xrb);
  // + 1 KB: the unused atoms of the last M tile read a few pixels past the halo
  p.x_stage_bytes = (p.x_tx_bytes[any_up ? 1 : 0] + 1024 + 1023) / 1024 * 1024;
  p.z_tx_bytes = (uint32_t)(16 * ((kx || thin) ? 8 * d->tx + 2 : 8 * d->tx) * p.zrb);
  if (kx)  // + one tile row: the discarded atom of tile 1 (filter row 3) reads one row past the 18
    p.x_stage_bytes = (p.x_tx_bytes[0] + (uint32_t)(p.pitch[0] * p.xrb) + 1023) / 1024 * 1024;
  if (thin)  // the M tile stacks 128 / cb filter rows: the discarded ones read up to tile row 15 + 128 / cb - 1
    p.x_stage_bytes = ((uint32_t)((16 + 128 / d->cb) * p.pitch[0] * p.xrb) + 1023) / 1024 * 1024;
  p.stage_bytes = p.x_stage_bytes + (p.z_tx_bytes + 1023) / 1024 * 1024;
  // The stage ring takes 128 KB, not all of shared memory: the weight gradient runs on the side stream beside the
  // BatchNorm-backward passes of the main stream (HBM-bound, 32 KB of shared memory per reduction CTA), and those
  // can only become resident on an SM whose weight-gradient CTA leaves them room.  With 224 KB (five stages of the
  // TX = 1 kernels) the two streams merely took turns on the SMs: step 10.75 -> 10.47 ms, c3 27.11 -> 26.74 ms,
  // c4 156.3 -> 153.4 ms; the kernel alone is as fast with two stages (MMR_WGRAD_SMEM_KB: A/B knob).
  static const int budget_kb = getenv("MMR_WGRAD_SMEM_KB") ? atoi(getenv("MMR_WGRAD_SMEM_KB")) : 128;
  int stages = (int)(((size_t)budget_kb * 1024) / p.stage_bytes);
  if (stages < 2) stages = 2;
  if (stages > kWhMaxStages) stages = kWhMaxStages;
  MMR_REQUIRE(stages >= 2, "wgrad: stage of %u bytes does not fit twice in shared memory", p.stage_bytes);
  p.stages = stages;
  p.n_split = d->n_split < 1 ? 1 : d->n_split;
  if (p.n_split > p.total_tiles) p.n_split = p.total_tiles;
  uint32_t cols = 32;
  while (cols < (uint32_t)(thin ? 3 * d->bn : (kx ? 2 * 3 * d->bn : p.A * d->bn))) cols <<= 1;
  p.tmem_cols = cols;
  p.partial = d->partial;
  p.dst = d->dst;
  p.dst_cout = d->dst_cout;
  p.dst_cin = d->dst_cin;
  size_t smem = (size_t)p.stages * p.stage_bytes + 256;
  if (smem < 116 * 1024) smem = 116 * 1024;  // one CTA per SM (TMEM budget)
  pl->smem_bytes = smem;

  const size_t maps_bytes = (maps.empty() ? 1 : maps.size()) * sizeof(CUtensorMap);
  cudaError_t e = cudaMalloc(&pl->dev_blob, maps_bytes);
  if (e != cudaSuccess) {
    delete pl;
    return fail("cudaMalloc(%zu) failed: %s", maps_bytes, cudaGetErrorString(e));
  }
  e = maps.empty() ? cudaSuccess : cudaMemcpy(pl->dev_blob, maps.data(), maps_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(pl->dev_blob);
    delete pl;
    return fail("cudaMemcpy failed: %s", cudaGetErrorString(e));
  }
  p.maps = reinterpret_cast<const CUtensorMap*>(pl->dev_blob);
  e = cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
  if (e != cudaSuccess) cudaGetLastError();
  if (kx) {
    e = d->tx == 1 ? (d->bn == 64 ? wh_kx_attr<1, 64>() : wh_kx_attr<1, 32>())
                   : (d->bn == 64 ? wh_kx_attr<2, 64>() : wh_kx_attr<2, 32>());
    if (e != cudaSuccess) cudaGetLastError();
  }
  if (thin) {
    e = wh_thin_dispatch(d->tx, d->cb, d->bn, nullptr, dim3(), 0, nullptr);
    if (e != cudaSuccess) cudaGetLastError();
  }
  *out_plan = pl;
  return 0;
}

extern "C" int mmr_wgrad_halo_plan_run(void* plan, int accumulate, mmr_stream_t stream) {
  MMR_REQUIRE(plan, "null plan");
  WhPlan* pl = reinterpret_cast<WhPlan*>(plan);
  const WhParams& p = pl->prm;
  dim3 grid(p.nchunks * p.n_ntiles, p.n_split);
  if (p.mode == 2) {
    wh_thin_dispatch(p.TX, p.cb, p.bn, &p, grid, pl->smem_bytes, as_stream(stream));
  } else if (p.mode == 1) {
    if (p.TX == 1 && p.bn == 64) mmr_launch((conv_wgrad_kx_kernel<1, 64>), grid, kWhThreads, pl->smem_bytes, as_stream(stream), p);
    else if (p.TX == 1) mmr_launch((conv_wgrad_kx_kernel<1, 32>), grid, kWhThreads, pl->smem_bytes, as_stream(stream), p);
    else if (p.bn == 64) mmr_launch((conv_wgrad_kx_kernel<2, 64>), grid, kWhThreads, pl->smem_bytes, as_stream(stream), p);
    else mmr_launch((conv_wgrad_kx_kernel<2, 32>), grid, kWhThreads, pl->smem_bytes, as_stream(stream), p);
  } else {
    mmr_launch((conv_wgrad_halo_kernel), grid, kWhThreads, pl->smem_bytes, as_stream(stream), p);
  }
  MMR_CUDA_CHECK(cudaGetLastError());
  const size_t total = (size_t)p.prow * (p.pcol / 4) * p.nchunks * p.n_ntiles;
  const int nl = p.n_split >= 8 ? 8 : (p.n_split >= 4 ? 4 : (p.n_split >= 2 ? 2 : 1));
  const int per = 256 / nl;
  int64_t blocks = (int64_t)((total + per - 1) / per);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  mmr_launch((wgrad_halo_reduce_kernel), (int)blocks, 256, 0, as_stream(stream), p, accumulate);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_wgrad_halo_plan_destroy(void* plan) {
  if (!plan) return 0;
  WhPlan* pl = reinterpret_cast<WhPlan*>(plan);
  if (pl->dev_blob) cudaFree(pl->dev_blob);
  delete pl;
  return 0;
}
