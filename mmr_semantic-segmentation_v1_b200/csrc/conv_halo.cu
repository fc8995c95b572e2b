// 3x3 / stride 1 / pad 1 convolution as an implicit GEMM on tcgen05, second kernel generation.
//
// What bounds the first generation (conv_gemm.cu) is L2 -> shared-memory traffic: every filter tap
// re-fetches its activation box and every tile re-fetches its weights.  Here one halo tile
// (18 rows x (8*TX+2) pixels x one channel chunk) is fetched ONCE per chunk and the nine taps
// are issued as nine shifted K-major views of it: with SWIZZLE_128B/64B/32B the hardware swizzle is
// a function of the absolute shared-memory address, so an operand may start at any row of the halo
// and use any row-group stride (scripts/probe/umma_shift_probe.cu, probe2.cu).  One weight slot is
// shared by the TX M-tiles of the macro tile.
//
//   M tile  = 16 rows x 8 pixels of the output (one 8-pixel row segment per UMMA 8-row group,
//             row-group stride = halo pitch), TX of them side by side form the macro tile;
//   N tile  = bn output channels;   K = chunks of cb channels (cb*2 = 128/64/32-byte rows) x 9 taps.
//
// A nearest-x2 upsampled source is replicated by the TMA engine itself: tensor maps with byte
// stride 0 on two extra dimensions (C, dupx, Wl, dupy, rows) land the upsampled halo in shared
// memory, so the concatenated / upsampled input of smp's DecoderBlock never exists in HBM.
//
// Warp roles (one persistent CTA per SM, 256 threads):
//   warp 0  halo producer (TMA)          warp 1  MMA issuer (one elected lane)
//   warp 2  TMEM allocator               warp 3  weight producer (TMA)
//   warps 4-7  epilogue: tcgen05.ld -> scale/bias/residual/ReLU -> bf16 -> swizzled staging tile in
//              shared memory -> TMA store; per-channel sum / sum of squares of the stored values
//              (BatchNorm batch statistics) are taken from the staging tile on the way.
// Accumulators are double-buffered in TMEM when 2*TX*bn <= 512 columns.
#include <stdlib.h>

#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace mmr {

constexpr int kHaloThreads = 256;
constexpr int kMaxChunks = 16;
constexpr int kMaxGroups = 16;
constexpr int kStatSlots = 8;

struct HaloChunk {
  int32_t map, map_edge, c0, up;
  const uint8_t* base;  // cp.async loader: channel c0 of pixel (0, 0, 0) of the source
  int32_t pxb, sw, sh;  //   bytes per source pixel, stored width / height (half the conv's when up)
};

// MMR_HALO_DBG bit 16: per-CTA timestamps of the last conv_halo launch (diagnostic; scripts/halo_trace.py):
// [0] entry, [1] after griddepcontrol.wait, [2] after setup, [3] MMA warp: first operands landed,
// [4] MMA warp: last commit issued, [5] epilogue warp 0: last item stored, [6] epilogue: finalisation done, [7] exit
#ifndef MMR_PREFETCH_DESC
#define MMR_PREFETCH_DESC 1
#endif
constexpr int kTraceSlots = 8;
__device__ unsigned long long g_halo_trace[256 * kTraceSlots];
#define MMR_TRACE(slot)                                                                              \
  do {                                                                                               \
    if ((p.dbg & 16) && blockIdx.x < 256) g_halo_trace[blockIdx.x * kTraceSlots + (slot)] = globaltimer_ns(); \
  } while (0)

struct HaloParams {
  const CUtensorMap* maps;  // device: [source maps ...][weight map][store-group maps ...]
  HaloChunk chunk[kMaxChunks];
  int32_t group_coff[kMaxGroups];
  int32_t group_ldc[kMaxGroups];
  __nv_bfloat16* group_ptr[kMaxGroups];
  int32_t group_pool[kMaxGroups];  // 1: 2x2 sum-pooled store into [N][H/2][W/2][ldc] (data gradient of a nearest-x2 source)
  int nchunks, wmap, smap0;
  int H, W, N;
  int TX, R, tiles_x, tiles_y, n_ntiles, total_items;  // R: output-row phases stacked along N (1, 2, 4)
  int rowbytes, KS;
  int bn, sg, gpn;  // store group = sg channels, gpn groups per N tile
  int tps, nslots;  // taps per weight slot, slots per chunk
  int halo_stages, w_slots, acc_bufs, out_stages;
  int direct;  // 1: bf16 tiles go to global memory straight from registers (narrow store groups)
  uint32_t halo_stage_bytes, w_slot_bytes, out_stage_bytes, w_tx_bytes;
  uint32_t halo_tx_bytes[2];
  int pitch[2];
  uint32_t tmem_cols;
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;
  int res_ldc, relu, out_mode;
  float* out_f32;
  int out_ldc, cout_total;
  double* stats;
  int stats_ld;
  MmrBnFinalize bnf;  // ticket == nullptr: not fused
  MmrBnBwdFused bb;   // z == nullptr: no fused BatchNorm backward sums
  MmrHeadMetric hm;   // head launches: argmax / confusion matrix in the epilogue (all NULL: plain logits)
  int dbg;  // diagnostics (MMR_HALO_DBG): 1 no MMA issue, 2 no epilogue work, 4 no halo TMA, 8 no weight TMA
  int epi_warps;  // 4, or 8: two warps per TMEM lane quarter, each taking 32 of the 64 channels of a store group
  int cpl;  // 1: halo tiles of 32- / 64-byte rows are gathered by warps 0 and 2 with cp.async
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_5d(void* dst, const void* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* map, const void* src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const void* map, const void* src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar(int nthreads = 128) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}
// the two warps (column halves) of one TMEM lane quarter
__device__ __forceinline__ void quarter_bar(int q) { asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory"); }

// Shared-memory matrix descriptor split in two words so that the issue loop only adds to the low
// one: hi = stride-dim offset | version 1 | swizzle; lo = (address >> 4) | LBO field 1.
__device__ __forceinline__ uint32_t desc_hi32(uint32_t sbo_bytes, uint32_t swz) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (swz << 29);
}
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                          uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\ntcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_lohi_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                              uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\ntcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
      : "memory");
}
// The MMA issuer's whole loop, specialised on (16-channel steps per tap, taps per weight slot, M tiles
// per macro tile): every descriptor offset inside a chunk is a compile-time multiple of a few
// uniform registers, so the elected lane spends 2-3 uniform instructions per tcgen05.mma.  With the
// generic loop the issue overhead (~190 clk per MMA) bounded the thin layers (N = 16, K = 16).
template <int KS, int TPS, int TX>
__device__ __forceinline__ void mma_warp_loop(const HaloParams& p, uint32_t tmem_base, uint32_t halo0, uint32_t w0,
                                              uint64_t* halo_full, uint64_t* halo_empty, uint64_t* w_full,
                                              uint64_t* w_empty, uint64_t* tmem_full, uint64_t* tmem_empty) {
  const uint32_t idesc = make_idesc_bf16(128, p.bn, 0, 0);
  const uint32_t rb = (uint32_t)p.rowbytes;
  const uint32_t swz = swizzle_code(p.rowbytes);
  const uint32_t hiB = desc_hi32(8 * rb, swz);
  const uint32_t tile_step = (8 * rb) >> 4;  // next M-tile of the macro tile: 8 pixels to the right
  const uint32_t px_step = rb >> 4;
  const uint32_t tap_step = ((uint32_t)p.bn * rb) >> 4;
  const uint32_t bn = (uint32_t)p.bn;
  int hs = 0, ws = 0, it = 0;
  uint32_t hph = 0, wph = 0;
  for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
    const int buf = it % p.acc_bufs;
    const uint32_t par = (uint32_t)(it / p.acc_bufs) & 1u;
    mbar_wait(&tmem_empty[buf], par ^ 1);
    tc_fence_after();
    const uint32_t d0 = tmem_base + (uint32_t)(buf * TX) * bn;
    for (int c = 0; c < p.nchunks; ++c) {
      const int up = p.chunk[c].up;
      const uint32_t pitch = (uint32_t)p.pitch[up];
      const uint32_t hiA = desc_hi32(pitch * rb, swz);
      const uint32_t row_step = (pitch * rb) >> 4;
      mbar_wait(&halo_full[hs], hph);
      if (p.cpl) fence_proxy_async_smem();  // cp.async wrote the tile through the generic proxy
      tc_fence_after();
      if (it == 0 && c == 0 && (threadIdx.x & 31) == 0) MMR_TRACE(3);
      const uint32_t a_stage = ((halo0 + (uint32_t)hs * p.halo_stage_bytes + (up ? rb : 0u)) >> 4) | 0x10000u;
#pragma unroll
      for (int s = 0; s < 9 / TPS; ++s) {
        mbar_wait(&w_full[ws], wph);
        tc_fence_after();
        const uint32_t b_slot = ((w0 + (uint32_t)ws * p.w_slot_bytes) >> 4) | 0x10000u;
        if (!(p.dbg & 1) && elect_one()) {
#pragma unroll
          for (int tt = 0; tt < TPS; ++tt) {
            const int tap = s * TPS + tt;  // compile-time after unrolling
            const uint32_t a_tap = a_stage + (uint32_t)(tap / 3) * row_step + (uint32_t)(tap % 3) * px_step;
            const uint32_t b_tap = b_slot + (uint32_t)tt * tap_step;
#pragma unroll
            for (int i = 0; i < TX; ++i) {
              const uint32_t d = d0 + (uint32_t)i * bn;
              const uint32_t a = a_tap + (uint32_t)i * tile_step;
              if (tap == 0)
                umma_lohi(d, a, hiA, b_tap, hiB, idesc, (uint32_t)(c != 0));
              else
                umma_lohi_acc(d, a, hiA, b_tap, hiB, idesc);
#pragma unroll
              for (int k = 1; k < KS; ++k) umma_lohi_acc(d, a + 2 * k, hiA, b_tap + 2 * k, hiB, idesc);
            }
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&w_empty[ws]);
        if (++ws == p.w_slots) {
          ws = 0;
          wph ^= 1;
        }
      }
      if (elect_one()) umma_commit(&halo_empty[hs]);
      if (++hs == p.halo_stages) {
        hs = 0;
        hph ^= 1;
      }
    }
    if (elect_one()) umma_commit(&tmem_full[buf]);
    __syncwarp();
  }
}

// Row-phase stacking (R = 2 or 4).  An M tile takes every R-th output row (row-group stride of the A
// descriptor = R halo rows), so the A view at halo row shift ty serves the output rows of phase
// r = ty-1, ty, ty+1 (filter rows ky = ty - r + 1) at once: ONE tcgen05.mma of N = (#phases) * bn
// against the weight rows [ky = 2, 1, 0] of filter column kx, which the slot holds in exactly that
// order, accumulates into the adjacent column blocks of those phases.  The 4 KB A operand -- what bounds
// an N = 64 MMA at the shared-memory port -- is read once for up to three phases: (R+2)*3 MMAs of average
// N = 3R/(R+2) * bn replace 9R MMAs of N = bn.  A phase is first touched at ty = r-1: at chunk 0,
// filter column 0, K step 0 the new phase gets its own non-accumulating MMA.
template <int KS, int TX, int R>
__device__ __forceinline__ void mma_warp_loop_r(const HaloParams& p, uint32_t tmem_base, uint32_t halo0, uint32_t w0,
                                                uint64_t* halo_full, uint64_t* halo_empty, uint64_t* w_full,
                                                uint64_t* w_empty, uint64_t* tmem_full, uint64_t* tmem_empty) {
  const uint32_t bn = (uint32_t)p.bn;
  const uint32_t idesc[4] = {0u, make_idesc_bf16(128, p.bn, 0, 0), make_idesc_bf16(128, 2 * p.bn, 0, 0),
                             make_idesc_bf16(128, 3 * p.bn, 0, 0)};
  const uint32_t rb = (uint32_t)p.rowbytes;
  const uint32_t swz = swizzle_code(p.rowbytes);
  const uint32_t hiB = desc_hi32(8 * rb, swz);
  const uint32_t tile_step = (8 * rb) >> 4;
  const uint32_t px_step = rb >> 4;
  const uint32_t ky_step = (bn * rb) >> 4;  // next filter row of the slot ([ky = 2, 1, 0][bn rows])
  int hs = 0, ws = 0, it = 0;
  uint32_t hph = 0, wph = 0;
  for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
    const int buf = it % p.acc_bufs;
    const uint32_t par = (uint32_t)(it / p.acc_bufs) & 1u;
    mbar_wait(&tmem_empty[buf], par ^ 1);
    tc_fence_after();
    const uint32_t d0 = tmem_base + (uint32_t)(buf * TX * R) * bn;
    for (int c = 0; c < p.nchunks; ++c) {
      const int up = p.chunk[c].up;
      const uint32_t pitch = (uint32_t)p.pitch[up];
      const uint32_t hiA = desc_hi32((uint32_t)R * pitch * rb, swz);
      const uint32_t row_step = (pitch * rb) >> 4;
      mbar_wait(&halo_full[hs], hph);
      if (p.cpl) fence_proxy_async_smem();  // cp.async wrote the tile through the generic proxy
      tc_fence_after();
      if (it == 0 && c == 0 && (threadIdx.x & 31) == 0) MMR_TRACE(3);
      const uint32_t a_stage = ((halo0 + (uint32_t)hs * p.halo_stage_bytes + (up ? rb : 0u)) >> 4) | 0x10000u;
#pragma unroll
      for (int s = 0; s < 3; ++s) {  // filter column kx = s: one weight slot
        mbar_wait(&w_full[ws], wph);
        tc_fence_after();
        const uint32_t b_slot = ((w0 + (uint32_t)ws * p.w_slot_bytes) >> 4) | 0x10000u;
        if (!(p.dbg & 1) && elect_one()) {
#pragma unroll
          for (int ty = -1; ty <= R; ++ty) {
            const int r_lo = ty - 1 < 0 ? 0 : ty - 1;
            const int r_hi = ty + 1 > R - 1 ? R - 1 : ty + 1;
            const int nph = r_hi - r_lo + 1;
            const int nold = (ty > R - 1 ? R - 1 : ty) - r_lo + 1;  // phases r <= ty: already touched
            const bool has_new = ty + 1 <= R - 1;
            const uint32_t a_ty = a_stage + (uint32_t)(ty + 1) * row_step + (uint32_t)s * px_step;
            const uint32_t b_ty = b_slot + (uint32_t)(1 - ty + r_lo) * ky_step;
#pragma unroll
            for (int i = 0; i < TX; ++i) {
              const uint32_t d = d0 + (uint32_t)(i * R + r_lo) * bn;
              const uint32_t a = a_ty + (uint32_t)i * tile_step;
              if (s == 0 && c == 0) {
                if (nold > 0) umma_lohi_acc(d, a, hiA, b_ty, hiB, idesc[nold > 0 ? nold : 1]);
                if (has_new)
                  umma_lohi(d + (uint32_t)(nold > 0 ? nold : 0) * bn, a, hiA,
                            b_ty + (uint32_t)(nold > 0 ? nold : 0) * ky_step, hiB, idesc[1], 0u);
              } else {
                umma_lohi_acc(d, a, hiA, b_ty, hiB, idesc[nph]);
              }
#pragma unroll
              for (int k = 1; k < KS; ++k) umma_lohi_acc(d, a + 2 * k, hiA, b_ty + 2 * k, hiB, idesc[nph]);
            }
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&w_empty[ws]);
        if (++ws == p.w_slots) {
          ws = 0;
          wph ^= 1;
        }
      }
      if (elect_one()) umma_commit(&halo_empty[hs]);
      if (++hs == p.halo_stages) {
        hs = 0;
        hph ^= 1;
      }
    }
    if (elect_one()) umma_commit(&tmem_full[buf]);
    __syncwarp();
  }
}

// XOR applied to the 16-byte chunk index of staging row r (matches TMA SWIZZLE_{128,64,32}B).
__device__ __forceinline__ uint32_t row_xor(uint32_t r, int rowbytes) {
  return rowbytes == 128 ? (r & 7u) : (rowbytes == 64 ? ((r >> 1) & 3u) : ((r >> 2) & 1u));
}

struct ItemCoord {
  int nt, x0, y0, n;
};
__device__ __forceinline__ ItemCoord decode_item(const HaloParams& p, int item) {
  ItemCoord c;
  c.nt = item % p.n_ntiles;
  int t = item / p.n_ntiles;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  c.n = t / p.tiles_y;
  c.x0 = tx * 8 * p.TX;
  c.y0 = ty * 16 * p.R;
  return c;
}

// Epilogue of every bf16 NHWC plan whose store groups use their default path (TMA store for 64-channel
// groups, straight-to-global for narrower ones).  SG, the statistics mode and "no scale / bias /
// residual / ReLU" (PLAIN: every training-time launch) are compile-time, so the per-group body is
// straight-line code.  Each epilogue warp owns the 32 accumulator rows of its TMEM lanes = 4 image rows
// x 8 pixels, stages them in its own 4 KB slice of the staging tile and issues its own TMA store:
// no CTA-wide barrier on the way, the four warps drift freely.
// STATS (needs one channel set per CTA: n_ntiles = gpn = 1): the BatchNorm batch statistics of the
// stored (bf16-rounded) values stay in registers for the whole launch -- one fp32 sum and sum of
// squares per channel and accumulator row -- and are reduced across rows once, by warp shuffles,
// when the CTA has finished its last item.
template <int N>
__device__ __forceinline__ void bulk_wait_read_n() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// One round (lane mask MK) of the halving butterfly that ends a statistics launch; CUR = channels a lane still
// holds.  Template recursion keeps every array index a compile-time constant (the arrays stay in registers).
template <int SG, int CUR, int MK>
__device__ __forceinline__ void colsum_butterfly(float (&s1)[SG], float (&s2)[SG], int lane) {
  const bool upper = (lane & MK) != 0;
  if constexpr (CUR > 1) {
    constexpr int half = CUR / 2;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send1 = upper ? s1[j] : s1[j + half], keep1 = upper ? s1[j + half] : s1[j];
      const float send2 = upper ? s2[j] : s2[j + half], keep2 = upper ? s2[j + half] : s2[j];
      s1[j] = keep1 + __shfl_xor_sync(0xffffffffu, send1, MK);
      s2[j] = keep2 + __shfl_xor_sync(0xffffffffu, send2, MK);
    }
    if constexpr (MK > 1) colsum_butterfly<SG, half, MK / 2>(s1, s2, lane);
  } else {
    s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], MK);
    s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], MK);
    if constexpr (MK > 1) colsum_butterfly<SG, 1, MK / 2>(s1, s2, lane);
  }
}

// STATS: 0 none; 1 forward statistics (sum, sum of squares of the stored values); 2 BatchNorm-backward sums of a
// data-gradient launch (MmrBnBwdFused): sum g and sum g*z with g = (z*msc + msh > 0) ? dx : 0, msc / msh read from
// the `bb_affine` staging in shared memory.

// Data gradient of a nearest-x2 source (smp DecoderBlock's F.interpolate): its gradient is the 2x2 sum of the
// conv-resolution gradient.  Instead of storing that 4x larger tensor for the BatchNorm-backward reduction to
// pool (write + read of 1.1 GB per step over the ten upsampled sources of U-Net++), the epilogue pools in registers:
// the two image rows of a window are two row phases of the same accumulator lane (rph >= 2) or lanes 8 apart
// (rph = 1), the two pixels are neighbouring lanes; fp32 sum, one bf16 rounding, 16-byte stores into the
// low-resolution tensor [N][H/2][W/2][ldc].
template <int SG>
__device__ __forceinline__ void epi_pool_group(const HaloParams& p, uint32_t tmem_base, uint64_t* tmem_empty, int q,
                                               int lane, const ItemCoord& ic, int buf, int g, int gi, bool last_group) {
  constexpr int CH = SG > 32 ? 32 : SG;
  const int m = q * 32 + lane, h = m >> 3, w = m & 7;
  __nv_bfloat16* const gptr = p.group_ptr[gi];
  const int gldc = p.group_ldc[gi], gcoff = p.group_coff[gi];
  const int Hl = p.H >> 1, Wl = p.W >> 1;
  const int R = p.R;
  const int npair = R >= 2 ? R / 2 : 1;
  for (int i = 0; i < p.TX; ++i) {
    for (int rp = 0; rp < npair; ++rp) {
      const int ir0 = i * R + (R >= 2 ? 2 * rp : 0);
      const bool last = last_group && i == p.TX - 1 && rp == npair - 1;
      const int y = ic.y0 + R * h + (R >= 2 ? 2 * rp : 0), x = ic.x0 + 8 * i + w;
      const bool keep = !(w & 1) && (R >= 2 || !((lane >> 3) & 1)) && y < p.H && x < p.W;
      __nv_bfloat16* dst = gptr + (((size_t)ic.n * Hl + (y >> 1)) * Wl + (x >> 1)) * gldc + gcoff;
#pragma unroll
      for (int c0 = 0; c0 < SG; c0 += CH) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                               (uint32_t)((buf * p.TX * R + ir0) * p.bn + g * SG + c0);
        uint32_t a[CH], b[CH];
#pragma unroll
        for (int k = 0; k < CH; k += 16) tmem_ld16(taddr + k, reinterpret_cast<uint32_t(&)[16]>(a[k]));
        if (R >= 2) {
#pragma unroll
          for (int k = 0; k < CH; k += 16) tmem_ld16(taddr + p.bn + k, reinterpret_cast<uint32_t(&)[16]>(b[k]));
        }
        tmem_ld_wait();
        if (last && c0 + CH >= SG) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
        float v[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          v[j] = __uint_as_float(a[j]);
          if (R >= 2)
            v[j] += __uint_as_float(b[j]);
          else
            v[j] += __shfl_xor_sync(0xffffffffu, v[j], 8);
          v[j] += __shfl_xor_sync(0xffffffffu, v[j], 1);
        }
        if (keep && !(p.dbg & 2)) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
#pragma unroll
          for (int k = 0; k < CH / 8; ++k)
            d4[k] = make_uint4(pack_bf16x2(v[8 * k], v[8 * k + 1]), pack_bf16x2(v[8 * k + 2], v[8 * k + 3]),
                               pack_bf16x2(v[8 * k + 4], v[8 * k + 5]), pack_bf16x2(v[8 * k + 6], v[8 * k + 7]));
        }
      }
    }
  }
}

// HV = 2 (SG = 64 only): eight epilogue warps, warp (q, hf) takes the channels [32 hf, 32 hf + 32) of the rows of TMEM
// lane quarter q.  The statistics epilogues were as slow as the MMAs of an item (4 warps, one per scheduler, 128
// accumulator registers each: a 9-10 us tail after the last MMA of every launch); with two warps per scheduler and
// half the columns each they run ahead of the MMA warp again.  The two warps of a quarter fill one staging slice and
// meet at a named barrier before warp hf = 0 issues its TMA store.
template <int SG, int STATS, bool PLAIN, int HV = 1>
__device__ __forceinline__ void epi_fast(const HaloParams& p, uint32_t tmem_base, uint8_t* out_base,
                                         uint64_t* tmem_full, uint64_t* tmem_empty, int q, int lane,
                                         const float* bb_affine, int hf = 0) {
  static_assert(HV == 1 || SG == 64, "column halves only for 64-channel store groups");
  constexpr bool STAGED = SG == 64;
  constexpr int WC = SG / HV;                       // channels of this warp
  constexpr int CH = (STATS && WC > 32) ? 32 : WC;  // accumulator columns per TMEM round trip
  constexpr int orb = SG * 2;                       // staging row bytes
  const int m = q * 32 + lane;                      // accumulator row = pixel (h, w) of the M tile
  const int h = m >> 3, w = m & 7;
  const int cw = hf * WC;                           // first channel of this warp inside the store group
  const uint32_t xr = row_xor((uint32_t)m, orb);
  float s1[STATS ? WC : 1], s2[STATS ? WC : 1];
#pragma unroll
  for (int j = 0; j < (STATS ? WC : 1); ++j) s1[j] = s2[j] = 0.f;
  int it = 0, gcount = 0;
  for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
    const ItemCoord ic = decode_item(p, item);
    const int buf = it % p.acc_bufs;
    const uint32_t par = (uint32_t)(it / p.acc_bufs) & 1u;
    if (STATS == 2) {
      // the z rows this thread will need for the item go to L2 while the MMAs of the item are still running
      for (int ir = 0; ir < p.TX * p.R; ++ir) {
        const int x = ic.x0 + 8 * (ir / p.R) + w, y = ic.y0 + p.R * h + ir % p.R;
        if (y < p.H && x < p.W) {
          const __nv_bfloat16* zr = reinterpret_cast<const __nv_bfloat16*>(p.bb.z) +
                                    (((size_t)ic.n * p.H + y) * p.W + x) * SG;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(zr + cw));
          if (WC == 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(zr + 32));
        }
      }
    }
    mbar_wait(&tmem_full[buf], par);
    tc_fence_after();
    if (p.dbg & 2) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      continue;
    }
    for (int g = 0; g < p.gpn; ++g) {
      const int ch0 = ic.nt * p.bn + g * SG;
      const int gi = ic.nt * p.gpn + g;
      __nv_bfloat16* const gptr = p.group_ptr[gi];
      const int gldc = p.group_ldc[gi], gcoff = p.group_coff[gi];
      if (STATS == 0 && PLAIN && HV == 1 && p.group_pool[gi]) {
        epi_pool_group<SG>(p, tmem_base, tmem_empty, q, lane, ic, buf, g, gi, g == p.gpn - 1);
        continue;
      }
      for (int ir = 0; ir < p.TX * p.R; ++ir, ++gcount) {
        const int i = ir / p.R, r = ir % p.R;   // M tile, output-row phase
        const int x = ic.x0 + 8 * i + w;
        const int y = ic.y0 + p.R * h + r;
        const bool valid = y < p.H && x < p.W;
        const size_t pix = ((size_t)ic.n * p.H + y) * p.W + x;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                               (uint32_t)((buf * p.TX * p.R + ir) * p.bn + g * SG + cw);
        uint8_t* stage = out_base + (size_t)(gcount % p.out_stages) * p.out_stage_bytes;
        if (STAGED && gcount >= p.out_stages) {
          // this warp's slice of the staging buffer is free once its store from out_stages groups
          // ago has been read out of shared memory
          if (lane == 0 && hf == 0) {
            if (p.out_stages == 1) bulk_wait_read_n<0>(); else bulk_wait_read_n<1>();
          }
          __syncwarp();
        }
        if (HV == 2) quarter_bar(q);   // the slice is free for both column halves
        const bool last = (g == p.gpn - 1) && (ir == p.TX * p.R - 1);
#pragma unroll
        for (int c0 = 0; c0 < WC; c0 += CH) {
          uint32_t r[CH];
          uint4 rz[STATS == 2 ? CH / 8 : 1];
          if (STATS == 2) {   // the unit's z at the same pixel: in flight while the accumulators are read
            const uint4* zp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bb.z) +
                                                             pix * SG + cw + c0);
#pragma unroll
            for (int k = 0; k < CH / 8; ++k) rz[k] = valid ? __ldg(zp + k) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int k = 0; k < CH; k += 16) tmem_ld16(taddr + c0 + k, reinterpret_cast<uint32_t(&)[16]>(r[k]));
          tmem_ld_wait();
          if (last && c0 + CH >= WC) {
            // every TMEM read of this item is done: hand the accumulators back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
          }
          float v[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
          if (!PLAIN) {
            if (bb_affine) {   // one channel set per CTA: scale / shift staged in shared memory (broadcast reads)
#pragma unroll
              for (int k = 0; k < CH / 4; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(bb_affine + cw + c0 + 4 * k);
                const float4 b = *reinterpret_cast<const float4*>(bb_affine + SG + cw + c0 + 4 * k);
                v[4 * k] = fmaf(v[4 * k], a.x, b.x);
                v[4 * k + 1] = fmaf(v[4 * k + 1], a.y, b.y);
                v[4 * k + 2] = fmaf(v[4 * k + 2], a.z, b.z);
                v[4 * k + 3] = fmaf(v[4 * k + 3], a.w, b.w);
              }
            } else {
              if (p.scale) {
#pragma unroll
                for (int j = 0; j < CH; ++j) v[j] *= __ldg(p.scale + min(ch0 + cw + c0 + j, p.cout_total - 1));
              }
              if (p.bias) {
#pragma unroll
                for (int j = 0; j < CH; ++j) v[j] += __ldg(p.bias + min(ch0 + cw + c0 + j, p.cout_total - 1));
              }
            }
            if (p.residual && valid) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.res_ldc + ch0 + cw + c0);
#pragma unroll
              for (int k = 0; k < CH / 8; ++k) {
                const uint4 rv = __ldg(rp + k);
                const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = unpack_bf16x2(rr[j]);
                  v[8 * k + 2 * j] += f.x;
                  v[8 * k + 2 * j + 1] += f.y;
                }
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
            }
          }
          uint32_t o[CH / 2];
#pragma unroll
          for (int j = 0; j < CH / 2; ++j) o[j] = valid ? pack_bf16x2(v[2 * j], v[2 * j + 1]) : 0u;
          if (STATS == 1) {
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) {
              const float f0 = __uint_as_float(o[j] << 16), f1 = __uint_as_float(o[j] & 0xffff0000u);
              s1[c0 + 2 * j] += f0;
              s2[c0 + 2 * j] = fmaf(f0, f0, s2[c0 + 2 * j]);
              s1[c0 + 2 * j + 1] += f1;
              s2[c0 + 2 * j + 1] = fmaf(f1, f1, s2[c0 + 2 * j + 1]);
            }
          }
          if (STATS == 2) {
#pragma unroll
            for (int k = 0; k < CH / 8; ++k) {
              const uint32_t zw[4] = {rz[k].x, rz[k].y, rz[k].z, rz[k].w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * k + jj, c = c0 + 2 * j;
                const float2 a = *reinterpret_cast<const float2*>(bb_affine + cw + c);        // mask scale
                const float2 b = *reinterpret_cast<const float2*>(bb_affine + SG + cw + c);   // mask shift
                const float z0 = __uint_as_float(zw[jj] << 16), z1 = __uint_as_float(zw[jj] & 0xffff0000u);
                const float d0 = __uint_as_float(o[j] << 16), d1 = __uint_as_float(o[j] & 0xffff0000u);
                const float g0 = fmaf(z0, a.x, b.x) > 0.f ? d0 : 0.f, g1 = fmaf(z1, a.y, b.y) > 0.f ? d1 : 0.f;
                s1[c] += g0;
                s2[c] = fmaf(g0, z0, s2[c]);
                s1[c + 1] += g1;
                s2[c + 1] = fmaf(g1, z1, s2[c + 1]);
              }
            }
          }
          if (STAGED) {
            uint8_t* rowp = stage + (size_t)m * orb;
#pragma unroll
            for (int k = 0; k < CH / 8; ++k)
              *reinterpret_cast<uint4*>(rowp + ((((uint32_t)((cw + c0) >> 3) + k) ^ xr) << 4)) =
                  make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
          } else if (valid) {
            __nv_bfloat16* dptr = gptr + pix * gldc + gcoff + cw + c0;
            if (((gldc | gcoff) & 15) == 0 && (reinterpret_cast<uintptr_t>(gptr) & 31) == 0) {
              // whole 32-byte sectors per store (STG.256): two 16-byte halves per sector cost the 16-channel layers
              // at 512^2 a third of their epilogue time
#pragma unroll
              for (int k = 0; k < CH / 16; ++k)
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dptr + 16 * k), "r"(o[8 * k]),
                             "r"(o[8 * k + 1]), "r"(o[8 * k + 2]), "r"(o[8 * k + 3]), "r"(o[8 * k + 4]), "r"(o[8 * k + 5]),
                             "r"(o[8 * k + 6]), "r"(o[8 * k + 7])
                             : "memory");
            } else {
              uint4* dst = reinterpret_cast<uint4*>(dptr);
#pragma unroll
              for (int k = 0; k < CH / 8; ++k) dst[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
            }
          }
        }
        if (STAGED) {
          fence_proxy_async_smem();
          if (HV == 2) quarter_bar(q); else __syncwarp();   // both column halves of the slice are written
          if (lane == 0 && hf == 0) {
            if (p.R == 1)
              tma_store_4d(&p.maps[p.smap0 + gi], stage + (size_t)q * 32 * orb, gcoff, ic.x0 + 8 * i, ic.y0 + 4 * q,
                           ic.n);
            else  // (C, W, R, H/R, N): the warp's four tile rows are R image rows apart
              tma_store_5d(&p.maps[p.smap0 + gi], stage + (size_t)q * 32 * orb, gcoff, ic.x0 + 8 * i, r,
                           ic.y0 / p.R + 4 * q, ic.n);
            bulk_commit();
          }
        }
      }
    }
  }
  if (STAGED && lane == 0 && hf == 0) bulk_wait0();
  if constexpr (STATS != 0) {
    // column sums over the 32 accumulator rows of this warp by a halving butterfly: in the round with lane
    // mask mk a lane keeps one half of its channels and hands the other half to its partner, so the SG sums
    // cost SG - SG/32 shuffles per array (62 at SG = 64) instead of 5 * SG (the plain per-channel reduction
    // ran as 640 latency-bound shuffles: 8 us at the end of every launch).  Afterwards lane l owns the
    // channels [l * SG / 32, (l + 1) * SG / 32) (SG = 16: lanes 2c and 2c + 1 both hold channel c).
    double* slot = (STATS == 2 ? p.bb.slots : p.stats) + (size_t)(blockIdx.x & (kStatSlots - 1)) * 2 * p.stats_ld;
    colsum_butterfly<WC, WC, 16>(s1, s2, lane);
    constexpr int PER = WC >= 32 ? WC / 32 : 1;
    const int cbase = cw + (WC >= 32 ? lane * PER : (lane >> 1));
    if (WC >= 32 || !(lane & 1)) {
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        if (cbase + j < p.cout_total) {
          atomicAdd(slot + cbase + j, (double)s1[j]);
          atomicAdd(slot + p.stats_ld + cbase + j, (double)s2[j]);
        }
      }
    }
  }
}

// Epilogue of the segmentation head: up to 16 classes, bias only, fp32 NCHW logits (what `model(img)`
// returns).  One TMEM round trip per M tile, the bias in registers, one predicated store per class: a
// warp's 32 pixels are 4 image rows x 8 consecutive pixels, i.e. four full 32-byte sectors per class plane.
// With a head metric (MmrHeadMetric) the same registers feed torch.argmax's rule and the confusion matrix:
// `hist` is a per-CTA [classes][classes] table of 32-bit counters in shared memory, flushed to the 64-bit
// global matrix of image n whenever the CTA moves on to another image (items are ordered image-major, so
// that is every few items) -- the logits of an eval step then never exist in HBM.
__device__ __forceinline__ void epi_head_f32(const HaloParams& p, uint32_t tmem_base, uint64_t* tmem_full,
                                             uint64_t* tmem_empty, int q, int lane, unsigned int* hist) {
  const int m = q * 32 + lane;
  const int h = m >> 3, w = m & 7;
  const size_t hw = (size_t)p.H * p.W;
  const int C = p.cout_total;
  const bool count = p.hm.confusion != nullptr && p.hm.labels != nullptr;
  float bias[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) bias[j] = (p.bias && j < C) ? __ldg(p.bias + j) : 0.f;
  if (count) {
    for (int k = m; k < C * C; k += 128) hist[k] = 0u;
    epi_bar();
  }
  int it = 0, cur_n = -1;
  auto flush = [&](int n) {   // all four epilogue warps: counters of image n -> global, then cleared
    epi_bar();
    for (int k = m; k < C * C; k += 128) {
      const unsigned int v = hist[k];
      if (v) atomicAdd(p.hm.confusion + (size_t)n * C * C + k, (unsigned long long)v);
      hist[k] = 0u;
    }
    epi_bar();
  };
  for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
    const ItemCoord ic = decode_item(p, item);
    const int buf = it % p.acc_bufs;
    const uint32_t par = (uint32_t)(it / p.acc_bufs) & 1u;
    if (count && ic.n != cur_n) {
      if (cur_n >= 0) flush(cur_n);
      cur_n = ic.n;
    }
    mbar_wait(&tmem_full[buf], par);
    tc_fence_after();
    const int y = ic.y0 + h;
    for (int i = 0; i < p.TX; ++i) {
      const int x = ic.x0 + 8 * i + w;
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.TX + i) * p.bn), r);
      tmem_ld_wait();
      if (i == p.TX - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      }
      if (y < p.H && x < p.W && !(p.dbg & 2)) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + bias[j];
        const size_t px = (size_t)y * p.W + x;
        if (p.out_f32) {
          float* dst = p.out_f32 + (size_t)ic.n * p.out_ldc * hw + px;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < C) dst[(size_t)j * hw] = v[j];
        }
        if (p.hm.pred_out || count) {
          float best = v[0];
          int arg = 0;
#pragma unroll
          for (int j = 1; j < 16; ++j) {
            // torch.argmax: first maximal index; NaN counts as maximal (mmr_confusion_from_logits' rule)
            if (j < C && (v[j] > best || (v[j] != v[j] && best == best))) best = v[j], arg = j;
          }
          if (p.hm.pred_out) p.hm.pred_out[(size_t)ic.n * hw + px] = (uint8_t)arg;
          if (count) {
            const long long t = p.hm.labels_u8
                                    ? (long long)__ldg(reinterpret_cast<const uint8_t*>(p.hm.labels) + (size_t)ic.n * hw + px)
                                    : __ldg(reinterpret_cast<const long long*>(p.hm.labels) + (size_t)ic.n * hw + px);
            if (t >= 0 && t < C) atomicAdd(&hist[(int)t * C + arg], 1u);
          }
        }
      }
    }
  }
  if (count && cur_n >= 0) flush(cur_n);
}


// Halo producer for narrow rows (16 / 32 channels per chunk = 32- / 64-byte pixels).  TMA moves such a tile
// one pixel row at a time (~4 clk per row: 75 us of an 84 us launch on the 16-channel 512^2 layers), so here
// the 32 lanes of warp 0 gather it with 16-byte cp.async instead: each lane computes the source pixel (nearest
// x2: low-resolution pixel (y >> 1, x >> 1)), zero-fills what lies outside the image (src-size 0) and writes
// the 16-byte chunk where SWIZZLE_32B / 64B puts it (chunk index ^ address bits [7, 7 + log2(chunks))).
// A stage is handed to the MMA warp by cp.async.mbarrier.arrive (one arrival per lane, when that lane's copies
// have landed); the MMA warp issues the generic -> async proxy fence (tcgen05.mma reads shared memory through
// the async proxy) after its wait.
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
constexpr int kCpWarps = 2;   // warps 0 and 2 (the TMEM allocator has nothing else to do) gather alternate rows
constexpr int kCpMaxK = 5;    // 16-byte chunks of one tile row per lane: ceil(36 pixels * 4 chunks / 32)
// The loop has to stay near 5 instructions per LDGSTS: scripts/probe/ldgsts_rate_probe.cu (profiles/r02r_ldgsts_rate.txt)
// measures 47-54 clk per warp-wide LDGSTS.128 with 16-32 of them in flight per warp (11 / 19 / 24 B per clk and SM with
// 1 / 2 / 4 warps: the HBM roofline needs 4), but 135 clk with only 4-8 in flight -- a first version with ~20 ALU
// instructions per copy and a producer-side wait_group + proxy fence per stage ran at 120-300 clk.  So everything
// that does not depend on the row is hoisted per item (source offset of the lane's chunks) or per launch (the
// swizzled destination offsets of the lane's chunks for the 8 possible phases of a row start: row * pitch * rowbytes
// modulo 1024), and a stage's copies are tracked by cp.async.mbarrier.arrive instead of a blocking wait.
template <int K>
__device__ __forceinline__ void halo_producer_cp_k(const HaloParams& p, uint8_t* halo_base, uint64_t* halo_full,
                                                   uint64_t* halo_empty, int lane, int wsel) {
  const int rows = 16 * p.R + 2;
  const uint32_t rb = (uint32_t)p.rowbytes;
  const int cpp_log = rb == 32 ? 1 : 2;       // 16-byte chunks per pixel: 2 or 4
  const uint32_t cmask = (1u << cpp_log) - 1u;
  const uint32_t jj = (uint32_t)lane & cmask;  // chunk index of every copy of this lane (32 % chunks per pixel == 0)
  int hs = 0;
  uint32_t hph = 0;
  for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
    const ItemCoord ic = decode_item(p, item);
    for (int c = 0; c < p.nchunks; ++c) {
      mbar_wait(&halo_empty[hs], hph ^ 1);
      if (!(p.dbg & 4)) {
        const HaloChunk ch = p.chunk[c];
        const int pitch = p.pitch[ch.up];
        const int per_row = pitch << cpp_log;
        const uint32_t dst0 = smem_u32(halo_base + (size_t)hs * p.halo_stage_bytes);
        const uint8_t* img = ch.base + (size_t)ic.n * ch.sh * ch.sw * ch.pxb;
        const int xs = ch.up ? ic.x0 - 2 : ic.x0 - 1;
        // per item: source byte offset of the lane's k-th chunk inside a row (-1: outside the image, -2: no chunk)
        uint32_t soff[K], ssz[K], px_off[K];   // source offset (clamped), 16 / 0 bytes to read, pixel offset in the tile
        bool have[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int e = lane + 32 * k;
          const int px = e >> cpp_log;
          int x = xs + px;
          const bool ok = x >= 0 && x < p.W;
          if (ch.up) x >>= 1;
          have[k] = e < per_row;
          soff[k] = ok ? (uint32_t)(x * ch.pxb) + jj * 16u : 0u;
          ssz[k] = ok ? 16u : 0u;
          px_off[k] = (uint32_t)px * rb + dst0;
        }
        const size_t row_stride = (size_t)ch.sw * ch.pxb;
        const uint32_t row_bytes = (uint32_t)pitch * rb;
#pragma unroll 2
        for (int r = wsel; r < rows; r += kCpWarps) {
          int y = ic.y0 - 1 + r;
          const bool row_ok = y >= 0 && y < p.H;
          if (ch.up) y >>= 1;
          const uint8_t* rowp = img + (row_ok ? (size_t)y * row_stride : 0);   // a valid address either way
          const uint32_t drow = (uint32_t)r * row_bytes;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            if (have[k]) {
              const uint32_t off = drow + px_off[k];     // dst0 is 1024-aligned: bits [7, 10) are the tile's own
              cp_async16_zfill(off + ((jj ^ ((off >> 7) & cmask)) << 4), rowp + soff[k], row_ok ? ssz[k] : 0u);
            }
          }
        }
      }
      // the lane's arrival fires when its copies of this stage have landed; the MMA warp then crosses to the
      // async proxy itself
      cp_async_mbar_arrive_noinc(&halo_full[hs]);
      if (++hs == p.halo_stages) {
        hs = 0;
        hph ^= 1;
      }
    }
  }
  cp_async_wait<0>();
}

__device__ __forceinline__ void halo_producer_cp(const HaloParams& p, uint8_t* halo_base, uint64_t* halo_full,
                                                 uint64_t* halo_empty, int lane, int wsel) {
  // chunks per tile row and lane, over every source of the launch
  const int cpp = p.rowbytes / 16;
  const int pitch = p.pitch[1] > p.pitch[0] ? p.pitch[1] : p.pitch[0];
  const int k = (pitch * cpp + 31) / 32;
  if (k <= 2) halo_producer_cp_k<2>(p, halo_base, halo_full, halo_empty, lane, wsel);
  else if (k == 3) halo_producer_cp_k<3>(p, halo_base, halo_full, halo_empty, lane, wsel);
  else halo_producer_cp_k<kCpMaxK>(p, halo_base, halo_full, halo_empty, lane, wsel);
}

// Epilogue family of a launch.  One __global__ instantiation per family: the register allocation and the code
// of the statistics / BatchNorm-backward / plain epilogues do not disturb each other (with all of them in one
// kernel the forward-statistics epilogue of a 64 -> 64 layer ran 20 % slower).
enum { kEpiOther = 0, kEpiStats = 1, kEpiPlain = 2, kEpiBnBwd = 3, kEpiAffine = 4, kEpiHead = 5, kEpiKinds = 6 };

template <int EK>
__device__ __forceinline__ void conv_halo_body(const HaloParams& p) {
  if (threadIdx.x == 0) MMR_TRACE(0);
#if MMR_PREFETCH_DESC
  // the tensor maps live in global memory (plan blob, written at build time): fetch the ones the first loads use
  // while the barriers / TMEM are being set up -- cold, a descriptor costs an HBM round trip in front of its first load
  if (!p.cpl && threadIdx.x >= 96 && threadIdx.x <= 96 + (unsigned)min(p.nchunks, 31)) {
    const int i = (int)threadIdx.x - 96;    // lanes of warp 3: one source chunk each, the last one the weights
    if (i == min(p.nchunks, 31)) {
      tma_prefetch_desc(&p.maps[p.wmap]);
    } else {
      tma_prefetch_desc(&p.maps[p.chunk[i].map]);
      if (p.chunk[i].up) tma_prefetch_desc(&p.maps[p.chunk[i].map_edge]);
    }
  }
#endif
  pdl_prologue_conv();   // late-trigger builds wait after the set-up below (pdl_setup_done)
  if (threadIdx.x == 0) MMR_TRACE(1);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* halo_base = smem;
  uint8_t* w_base = halo_base + (size_t)p.halo_stages * p.halo_stage_bytes;
  uint8_t* out_base = w_base + (size_t)p.w_slots * p.w_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_base + (size_t)p.out_stages * p.out_stage_bytes);
  uint64_t* halo_full = bars;
  uint64_t* halo_empty = bars + 4;
  uint64_t* w_full = bars + 8;
  uint64_t* w_empty = bars + 16;
  uint64_t* tmem_full = bars + 24;
  uint64_t* tmem_empty = bars + 26;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 28);

  // warp index through a shuffle: tells the compiler it is warp-uniform, which keeps the role
  // loops (descriptor arithmetic, barrier indices) on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();  // swizzled operands need the 1024-byte alignment
    for (int s = 0; s < p.halo_stages; ++s) {
      mbar_init(&halo_full[s], p.cpl ? 32 * kCpWarps : 1);
      mbar_init(&halo_empty[s], 1);
    }
    for (int s = 0; s < p.w_slots; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], p.epi_warps);
    }
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
  pdl_setup_done();      // nothing above reads or writes global memory
  if (threadIdx.x == 0) MMR_TRACE(2);

  if ((warp == 0 || warp == 2) && p.cpl) {
    halo_producer_cp(p, halo_base, halo_full, halo_empty, lane, warp >> 1);
  } else if (warp == 0) {
    // ---------------------------------------------------------------- halo producer
    int hs = 0;
    uint32_t hph = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const ItemCoord ic = decode_item(p, item);
      for (int c = 0; c < p.nchunks; ++c) {
        mbar_wait(&halo_empty[hs], hph ^ 1);
        if ((p.dbg & 4) && elect_one()) mbar_arrive(&halo_full[hs]);
        if (!(p.dbg & 4) && elect_one()) {
          const HaloChunk ch = p.chunk[c];
          uint8_t* dst = halo_base + (size_t)hs * p.halo_stage_bytes;
          if (!ch.up) {
            mbar_arrive_expect_tx(&halo_full[hs], p.halo_tx_bytes[0]);
            tma_load_4d(dst, &p.maps[ch.map], &halo_full[hs], ch.c0, ic.x0 - 1, ic.y0 - 1, ic.n);
          } else {
            // halo row 0 = upsampled row y0-1, rows 1..16R = y0..y0+16R-1, row 16R+1 = y0+16R;
            // halo column 0 = upsampled column x0-2 (even, so the pair replication lines up).
            const int xl = (ic.x0 >> 1) - 1, yl = ic.y0 >> 1;
            const uint32_t rowb = (uint32_t)p.pitch[1] * p.rowbytes;
            mbar_arrive_expect_tx(&halo_full[hs], p.halo_tx_bytes[1]);
            tma_load_5d(dst, &p.maps[ch.map_edge], &halo_full[hs], ch.c0, 0, xl, yl - 1, ic.n);
            tma_load_5d(dst + rowb, &p.maps[ch.map], &halo_full[hs], ch.c0, 0, xl, 0,
                        ic.n * (p.H >> 1) + yl);
            tma_load_5d(dst + (size_t)(16 * p.R + 1) * rowb, &p.maps[ch.map_edge], &halo_full[hs], ch.c0, 0, xl,
                        yl + 8 * p.R, ic.n);
          }
        }
        __syncwarp();
        if (++hs == p.halo_stages) {
          hs = 0;
          hph ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ---------------------------------------------------------------- weight producer
    int ws = 0;
    uint32_t wph = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int nt = item % p.n_ntiles;
      int row = nt * p.nchunks * 9 * p.bn;
      for (int c = 0; c < p.nchunks; ++c) {
        for (int s = 0; s < p.nslots; ++s) {
          mbar_wait(&w_empty[ws], wph ^ 1);
          if ((p.dbg & 8) && elect_one()) mbar_arrive(&w_full[ws]);
          if (!(p.dbg & 8) && elect_one()) {
            mbar_arrive_expect_tx(&w_full[ws], p.w_tx_bytes);
            tma_load_2d(w_base + (size_t)ws * p.w_slot_bytes, &p.maps[p.wmap], &w_full[ws], 0, row);
          }
          __syncwarp();
          row += p.tps * p.bn;
          if (++ws == p.w_slots) {
            ws = 0;
            wph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t halo0 = smem_u32(halo_base), w0 = smem_u32(w_base);
#define MMR_MMA_CASE(KS_, TPS_, TX_)                                                                        \
  if (p.KS == KS_ && p.tps == TPS_ && p.TX == TX_)                                                          \
    mma_warp_loop<KS_, TPS_, TX_>(p, tmem_base, halo0, w0, halo_full, halo_empty, w_full, w_empty, tmem_full, \
                                  tmem_empty);
#define MMR_MMA_CASES_TX(KS_, TPS_) MMR_MMA_CASE(KS_, TPS_, 1) MMR_MMA_CASE(KS_, TPS_, 2) MMR_MMA_CASE(KS_, TPS_, 4)
#define MMR_MMA_CASES_TPS(KS_) MMR_MMA_CASES_TX(KS_, 1) MMR_MMA_CASES_TX(KS_, 3) MMR_MMA_CASES_TX(KS_, 9)
    if (p.R == 1) {
      MMR_MMA_CASES_TPS(1)
      MMR_MMA_CASES_TPS(2)
      MMR_MMA_CASES_TPS(4)
    }
#define MMR_MMA_RCASE(KS_, TX_, R_)                                                                             \
  if (p.KS == KS_ && p.TX == TX_ && p.R == R_)                                                                  \
    mma_warp_loop_r<KS_, TX_, R_>(p, tmem_base, halo0, w0, halo_full, halo_empty, w_full, w_empty, tmem_full,   \
                                  tmem_empty);
#define MMR_MMA_RCASES(KS_) \
  MMR_MMA_RCASE(KS_, 1, 2) MMR_MMA_RCASE(KS_, 2, 2) MMR_MMA_RCASE(KS_, 4, 2) MMR_MMA_RCASE(KS_, 1, 4) MMR_MMA_RCASE(KS_, 2, 4)
    MMR_MMA_RCASES(1)
    MMR_MMA_RCASES(2)
    MMR_MMA_RCASES(4)
#undef MMR_MMA_RCASES
#undef MMR_MMA_RCASE
#undef MMR_MMA_CASES_TPS
#undef MMR_MMA_CASES_TX
#undef MMR_MMA_CASE
    if (lane == 0) MMR_TRACE(4);
    pdl_done();          // every MMA of this CTA is issued: the next kernel's launch overlaps the last epilogue
  } else if (warp >= 4 && warp < 4 + p.epi_warps) {
    // ---------------------------------------------------------------- epilogue
    const int q = (warp - 4) & 3;   // TMEM lane quarter (a warp may only read the lanes 32 (warp % 4) ...)
    const int hf = (warp - 4) >> 2; // column half (eight epilogue warps: statistics / BatchNorm-backward launches)
    const int m = q * 32 + lane;    // accumulator row = pixel (h, w) of the M tile
    const int et = (warp - 4) * 32 + lane, en = 32 * p.epi_warps;  // thread index / count of the epilogue group
    // fast path: default store mode of the group width, statistics (if any) of one channel set per CTA
    const bool plain = !p.scale && !p.bias && !p.residual && !p.relu;
    const bool fast = p.out_mode == MMR_OUT_BF16_NHWC && (p.direct != 0) == (p.sg < 64) &&
                      (p.stats == nullptr || (p.n_ntiles == 1 && p.gpn == 1 && plain));
    const bool head = p.out_mode == MMR_OUT_F32_NCHW && p.bn == 16 && p.n_ntiles == 1 && p.R == 1 && !p.scale &&
                      !p.residual && !p.relu;
    (void)fast;
    (void)head;
    if constexpr (EK == kEpiBnBwd) {
      // data gradient with the consumer unit's BatchNorm-backward sums: mask scale / shift staged in shared memory
      float* bb_affine = reinterpret_cast<float*>(bars + 32);
      for (int c = et; c < p.sg; c += en) {
        bb_affine[c] = __ldg(p.bb.mask_scale + c);
        bb_affine[p.sg + c] = __ldg(p.bb.mask_shift + c);
      }
      epi_bar(en);
      if (p.sg == 64) epi_fast<64, 2, true, 2>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, bb_affine, hf);
      if (p.sg == 32) epi_fast<32, 2, true>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, bb_affine);
      if (p.sg == 16) epi_fast<16, 2, true>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, bb_affine);
    } else if constexpr (EK == kEpiStats) {
      if (p.sg == 64) epi_fast<64, 1, true, 2>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, nullptr, hf);
      if (p.sg == 32) epi_fast<32, 1, true>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, nullptr);
      if (p.sg == 16) epi_fast<16, 1, true>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, nullptr);
    } else if constexpr (EK == kEpiPlain) {
      if (p.sg == 64) epi_fast<64, 0, true>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, nullptr);
      if (p.sg == 32) epi_fast<32, 0, true>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, nullptr);
      if (p.sg == 16) epi_fast<16, 0, true>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, nullptr);
    } else if constexpr (EK == kEpiHead) {
      epi_head_f32(p, tmem_base, tmem_full, tmem_empty, q, lane, reinterpret_cast<unsigned int*>(bars + 32));
    } else if constexpr (EK == kEpiAffine) {   // scale / bias / residual / ReLU epilogues (eval mode, biased convs)
      float* aff = nullptr;
      if (p.n_ntiles == 1 && p.gpn == 1) {     // the whole channel set is one store group: stage scale / shift once
        aff = reinterpret_cast<float*>(bars + 32);
        for (int c = et; c < p.sg; c += en) {
          const int cc = min(c, p.cout_total - 1);
          aff[c] = p.scale ? __ldg(p.scale + cc) : 1.f;
          aff[p.sg + c] = p.bias ? __ldg(p.bias + cc) : 0.f;
        }
        epi_bar(en);
      }
      if (p.sg == 64 && p.epi_warps == 8) epi_fast<64, 0, false, 2>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, aff, hf);
      if (p.sg == 64 && p.epi_warps == 4) epi_fast<64, 0, false>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, aff);
      if (p.sg == 32) epi_fast<32, 0, false>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, aff);
      if (p.sg == 16) epi_fast<16, 0, false>(p, tmem_base, out_base, tmem_full, tmem_empty, q, lane, aff);
    } else {
    const int h = m >> 3, w = m & 7;
    const int orb = p.sg * 2;  // staging row bytes
    const uint32_t xr = row_xor((uint32_t)m, orb);
    const int npairs = p.sg >> 1;
    const int pr = m % npairs, rg = m / npairs, nrg = 128 / npairs;
    const int slot = blockIdx.x & (kStatSlots - 1);
    // one channel set per CTA: keep the statistics in registers for the whole launch
    const bool persist = p.n_ntiles == 1 && p.gpn == 1;
    const bool staged = p.out_mode == MMR_OUT_BF16_NHWC && (!p.direct || p.stats != nullptr);
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
    auto flush_stats = [&](int ch0) {
      const int ch = ch0 + 2 * pr;
      if (ch < p.cout_total) {
        double* s = p.stats + (size_t)slot * 2 * p.stats_ld;
        atomicAdd(s + ch, (double)s1a);
        atomicAdd(s + p.stats_ld + ch, (double)s2a);
        if (ch + 1 < p.cout_total) {
          atomicAdd(s + ch + 1, (double)s1b);
          atomicAdd(s + p.stats_ld + ch + 1, (double)s2b);
        }
      }
      s1a = s1b = s2a = s2b = 0.f;
    };
    int it = 0, gcount = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      const ItemCoord ic = decode_item(p, item);
      const int buf = it % p.acc_bufs;
      const uint32_t par = (uint32_t)(it / p.acc_bufs) & 1u;
      mbar_wait(&tmem_full[buf], par);
      tc_fence_after();
      const int y = ic.y0 + h;
      for (int g = 0; g < p.gpn; ++g) {
        const int ch0 = ic.nt * p.bn + g * p.sg;
        for (int i = 0; i < p.TX; ++i) {
          const int x = ic.x0 + 8 * i + w;
          const bool valid = y < p.H && x < p.W;
          const size_t pix = ((size_t)ic.n * p.H + y) * p.W + x;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                 (uint32_t)(buf * p.TX * p.bn + i * p.bn + g * p.sg);
          uint8_t* stage = out_base;
          if (staged) stage += (size_t)(gcount % p.out_stages) * p.out_stage_bytes;
          if (staged && p.out_stages == 1 && gcount > 0) {
            if (m == 0 && !p.direct) bulk_wait_read0();
            epi_bar();
          }
          const bool last = (g == p.gpn - 1) && (i == p.TX - 1);
          for (int c0 = 0; c0 < p.sg; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait();
            if (last && c0 + 16 >= p.sg) {
              // every TMEM read of this item is done: hand the accumulators back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            }
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
            if (p.scale) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] *= __ldg(p.scale + min(ch0 + c0 + j, p.cout_total - 1));
            }
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += __ldg(p.bias + min(ch0 + c0 + j, p.cout_total - 1));
            }
            if (p.residual && valid) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.res_ldc + ch0 + c0);
              const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
              const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 f = unpack_bf16x2(rr[j]);
                v[2 * j] += f.x;
                v[2 * j + 1] += f.y;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.out_mode == MMR_OUT_BF16_NHWC) {
              uint4 o0, o1;
              if (valid) {
                o0.x = pack_bf16x2(v[0], v[1]);
                o0.y = pack_bf16x2(v[2], v[3]);
                o0.z = pack_bf16x2(v[4], v[5]);
                o0.w = pack_bf16x2(v[6], v[7]);
                o1.x = pack_bf16x2(v[8], v[9]);
                o1.y = pack_bf16x2(v[10], v[11]);
                o1.z = pack_bf16x2(v[12], v[13]);
                o1.w = pack_bf16x2(v[14], v[15]);
              } else {
                o0 = make_uint4(0, 0, 0, 0);
                o1 = o0;
              }
              if (staged) {
                uint8_t* rowp = stage + (size_t)m * orb;
                const uint32_t j0 = (uint32_t)c0 >> 3;
                *reinterpret_cast<uint4*>(rowp + ((j0 ^ xr) << 4)) = o0;
                *reinterpret_cast<uint4*>(rowp + (((j0 + 1) ^ xr) << 4)) = o1;
              }
              if (p.direct && valid) {
                const int gi = ic.nt * p.gpn + g;
                uint4* dst = reinterpret_cast<uint4*>(p.group_ptr[gi] + pix * p.group_ldc[gi] + p.group_coff[gi] + c0);
                dst[0] = o0;
                dst[1] = o1;
              }
            } else if (valid) {
              const size_t hw = (size_t)p.H * p.W;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int ch = ch0 + c0 + j;
                if (ch < p.cout_total)
                  p.out_f32[((size_t)ic.n * p.out_ldc + ch) * hw + (size_t)y * p.W + x] = v[j];
              }
            }
          }
          if (staged) {
            if (!p.direct) {
              if (p.out_stages > 1 && m == 0) bulk_wait_read0();  // the other staging buffer is free again
              fence_proxy_async_smem();
            }
            epi_bar();
            if (m == 0 && !p.direct) {
              const int gi = ic.nt * p.gpn + g;
              for (int qq = 0; qq < 4; ++qq)  // the store box is one warp's slice: 4 image rows
                tma_store_4d(&p.maps[p.smap0 + gi], stage + (size_t)qq * 32 * orb, p.group_coff[gi], ic.x0 + 8 * i,
                             ic.y0 + 4 * qq, ic.n);
              bulk_commit();
            }
            if (p.stats) {
              for (int k = 0; k < npairs; ++k) {
                const uint32_t r = (uint32_t)(rg + k * nrg);
                const uint32_t off = r * orb + ((((uint32_t)pr >> 2) ^ row_xor(r, orb)) << 4) + ((uint32_t)pr & 3u) * 4u;
                const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(stage + off));
                s1a += f.x;
                s1b += f.y;
                s2a += f.x * f.x;
                s2b += f.y * f.y;
              }
            }
            ++gcount;
          }
        }
        if (p.stats && !persist) flush_stats(ch0);
      }
    }
    if (p.stats && persist) flush_stats(0);
    }  // generic epilogue
    if (et == 0) MMR_TRACE(5);
    if (p.stats && p.bnf.ticket) {
      // fused BatchNorm finalisation: the last CTA to get here owns the complete sums
      uint32_t* flag = reinterpret_cast<uint32_t*>(bars + 30);
      __threadfence();
      epi_bar(en);
      if (et == 0) *flag = atomicAdd(p.bnf.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
      epi_bar(en);
      if (*flag) {
        __threadfence();
        const double P = (double)p.bnf.count;
        for (int c = et; c < p.cout_total; c += en) {
          double s1 = 0.0, s2 = 0.0;
          for (int sl = 0; sl < kStatSlots; ++sl) {
            double* a = p.stats + (size_t)sl * 2 * p.stats_ld + c;
            s1 += __ldcg(a);
            s2 += __ldcg(a + p.stats_ld);
            a[0] = 0.0;
            a[p.stats_ld] = 0.0;
          }
          const double mu = s1 / P;
          double var = s2 / P - mu * mu;
          if (var < 0.0) var = 0.0;
          const float is = (float)(1.0 / sqrt(var + (double)p.bnf.eps));
          const float ga = p.bnf.gamma ? p.bnf.gamma[c] : 1.f, be = p.bnf.beta ? p.bnf.beta[c] : 0.f;
          p.bnf.mean[c] = (float)mu;
          p.bnf.invstd[c] = is;
          p.bnf.scale[c] = ga * is;
          p.bnf.shift[c] = be - (float)mu * ga * is;
          if (p.bnf.running_mean) {
            const double unbiased = P > 1.0 ? var * P / (P - 1.0) : var;
            p.bnf.running_mean[c] = (1.f - p.bnf.momentum) * p.bnf.running_mean[c] + p.bnf.momentum * (float)mu;
            p.bnf.running_var[c] = (1.f - p.bnf.momentum) * p.bnf.running_var[c] + p.bnf.momentum * (float)unbiased;
          }
        }
        if (et == 0) {
          *p.bnf.ticket = 0u;
          if (p.bnf.num_batches_tracked) p.bnf.num_batches_tracked[0] += 1;
        }
      }
    }
    if (p.bb.z) {
      // fused BatchNorm-backward finalisation (the arithmetic of reduce_rows_kernel's last CTA)
      uint32_t* flag = reinterpret_cast<uint32_t*>(bars + 30);
      __threadfence();
      epi_bar(en);
      if (et == 0) *flag = atomicAdd(p.bb.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
      epi_bar(en);
      if (*flag) {
        __threadfence();
        const int Cn = p.cout_total;
        for (int c = et; c < Cn; c += en) {
          double s1 = 0.0, s2 = 0.0;
          for (int sl = 0; sl < kStatSlots; ++sl) {
            double* a = p.bb.slots + (size_t)sl * 2 * p.stats_ld + c;
            s1 += __ldcg(a);
            s2 += __ldcg(a + p.stats_ld);
            a[0] = 0.0;
            a[p.stats_ld] = 0.0;
          }
          s2 = (s2 - (double)p.bb.mean[c] * s1) * (double)p.bb.invstd[c];   // sum g*xhat
          if (p.bb.dgamma) p.bb.dgamma[c] = (p.bb.accumulate ? p.bb.dgamma[c] : 0.f) + (float)s2;
          if (p.bb.dbeta) p.bb.dbeta[c] = (p.bb.accumulate ? p.bb.dbeta[c] : 0.f) + (float)s1;
          const double gi = (double)(p.bb.gamma ? p.bb.gamma[c] : 1.f) * (double)p.bb.invstd[c];
          p.bb.coef[c] = (float)gi;
          p.bb.coef[Cn + c] = (float)(-gi * s2 / (double)p.bb.count);
          p.bb.coef[2 * Cn + c] = (float)(-gi * s1 / (double)p.bb.count);
        }
        if (et == 0) *p.bb.ticket = 0u;
      }
    }
    if (p.out_mode == MMR_OUT_BF16_NHWC && !p.direct && et == 0) bulk_wait0();
    if (et == 0) MMR_TRACE(6);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
  if (threadIdx.x == 0) MMR_TRACE(7);
}

// The statistics / BatchNorm-backward families run with eight epilogue warps (384 threads, 168 registers each).
constexpr int halo_threads(int ek) {
  return (ek == kEpiStats || ek == kEpiBnBwd || ek == kEpiAffine) ? kHaloThreads + 128 : kHaloThreads;
}

template <int EK>
__global__ void __launch_bounds__(halo_threads(EK), 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<EK>(p);
}

// ------------------------------------------------------------------ weight packing
// out[((nt*nchunks + c)*9 + tap)*bn + r][k], bf16.  mode 0 (fprop): N index = output channel,
// K index = concatenated input channel, filter tap as is.  mode 1 (dgrad): N index = input channel,
// K index = output channel, tap mirrored (the data gradient correlates dz with the flipped filter).
// layout 1 (row-phase stacking): slot position t holds filter column kx = t / 3, filter row ky = 2 - t % 3.
__device__ __forceinline__ int pack_tap(int t, int layout) { return layout ? (2 - t % 3) * 3 + t / 3 : t; }

__global__ void pack_weights_halo_kernel(const float* __restrict__ w, int O, int I, int mode, int cb, int bn,
                                         int n_ntiles, int nchunks, int layout, __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const int64_t total = (int64_t)n_ntiles * nchunks * 9 * bn * cb;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = idx;
    const int k = (int)(t % cb);
    t /= cb;
    const int r = (int)(t % bn);
    t /= bn;
    const int tap = pack_tap((int)(t % 9), layout);
    t /= 9;
    const int c = (int)(t % nchunks);
    const int nt = (int)(t / nchunks);
    const int nidx = nt * bn + r, kidx = c * cb + k;
    float v = 0.f;
    if (mode == 0) {
      if (nidx < O && kidx < I) v = w[((size_t)nidx * I + kidx) * 9 + tap];
    } else {
      if (kidx < O && nidx < I) v = w[((size_t)kidx * I + nidx) * 9 + (8 - tap)];
    }
    out[idx] = __float2bfloat16(v);
  }
}

// Every layer's packing in one launch.  A block owns kPackPerBlock consecutive (N tile, chunk, row, k) items
// of one job and finds its job by walking the job sizes (staged in shared memory); a thread takes eight
// consecutive k of one row: in the fprop layout those are 72 consecutive floats of the fp32 master (eighteen
// 16-byte loads), in the dgrad layout eight 36-byte records; either way nine 16-byte stores, one per tap plane
// (one thread per (row, k) with 2-byte stores took 96 us per step for 190 MB).
constexpr int kPackThreads = 128, kPackIters = 8;
constexpr int kPackPerBlock = kPackThreads * 8 * kPackIters;   // 8192 items: the job walk is paid once per 16 KB written
constexpr int kMaxPackJobs = 512;
__device__ __forceinline__ int64_t pack_items(const MmrPackJob& j) {
  return (int64_t)j.n_ntiles * j.nchunks * j.bn * j.cb;
}
__global__ void __launch_bounds__(kPackThreads)
pack_weights_halo_batch_kernel(const MmrPackJob* __restrict__ jobs, int njobs) {
  pdl_prologue();
  __shared__ int64_t totals[kMaxPackJobs];
  for (int i = threadIdx.x; i < njobs; i += blockDim.x) totals[i] = pack_items(jobs[i]);
  __syncthreads();
  int64_t b = blockIdx.x, total = 0;
  int ji = 0;
  for (; ji < njobs; ++ji) {
    total = totals[ji];
    const int64_t nb = (total + kPackPerBlock - 1) / kPackPerBlock;
    if (b < nb) break;
    b -= nb;
  }
  if (ji == njobs) return;
  const MmrPackJob j = jobs[ji];
  const float* __restrict__ w = j.w_oihw;
  __nv_bfloat16* __restrict__ out = reinterpret_cast<__nv_bfloat16*>(j.out);
  const uint32_t cb = (uint32_t)j.cb, bn = (uint32_t)j.bn, nch = (uint32_t)j.nchunks;
  // one thread and iteration = eight consecutive k of one (N tile, chunk, row): nine 16-byte stores, one per tap plane
  for (int it = 0; it < kPackIters; ++it) {
  const int64_t idx = b * kPackPerBlock + ((int64_t)it * kPackThreads + threadIdx.x) * 8;
  if (idx >= total) return;
  uint32_t t = (uint32_t)idx;
  const uint32_t k0 = t % cb;
  t /= cb;
  const uint32_t r = t % bn;
  t /= bn;
  const uint32_t c = t % nch, nt = t / nch;
  const int nidx = (int)(nt * bn + r), kidx0 = (int)(c * cb + k0);
  float v[8][9];
  if (j.mode == 0) {
    // fprop: the eight (output channel nidx, input channel kidx0 + q) records are 72 consecutive floats
    const float* src = w + ((size_t)nidx * j.I + kidx0) * 9;
    if (nidx < j.O && kidx0 + 8 <= j.I && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      float f[72];
#pragma unroll
      for (int q = 0; q < 18; ++q) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(src) + q);
        f[4 * q] = x.x, f[4 * q + 1] = x.y, f[4 * q + 2] = x.z, f[4 * q + 3] = x.w;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) v[q][tp] = f[q * 9 + tp];
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const bool live = nidx < j.O && kidx0 + q < j.I;
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) v[q][tp] = live ? __ldg(src + q * 9 + tp) : 0.f;
      }
    }
  } else {
    // dgrad: N index = input channel, K index = output channel, taps mirrored below
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const bool live = kidx0 + q < j.O && nidx < j.I;
      const float* src = w + ((size_t)(kidx0 + q) * j.I + nidx) * 9;
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) v[q][tp] = live ? __ldg(src + tp) : 0.f;
    }
  }
  // out[((nt*nchunks + c)*9 + slot)*bn + r][k], slot -> filter tap by layout, mirrored for dgrad
  __nv_bfloat16* dst = out + ((size_t)(nt * nch + c) * 9 * bn + r) * cb + k0;
  // walk the SOURCE taps (static register indices) and compute the slot each one goes to: the inverse of
  // pack_tap (layout 1: slot = kx * 3 + (2 - ky)), after the mirror for dgrad
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) {
    const int tap = j.mode == 0 ? tp : 8 - tp;
    const int slot = j.layout ? (tap % 3) * 3 + (2 - tap / 3) : tap;
    uint32_t pk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * q][tp], v[2 * q + 1][tp]);
      pk[q] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(dst + (size_t)slot * bn * cb) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
void* get_encode_tiled();  // common.cu

static int encode_generic(CUtensorMap* out, const void* ptr, int rank, const cuuint64_t* dims,
                          const cuuint64_t* strides, const cuuint32_t* box, int inner_bytes, const char* what) {
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(get_encode_tiled());
  MMR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  MMR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: base must be 16-byte aligned", what);
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                    : CU_TENSOR_MAP_SWIZZLE_32B;
  for (int i = 0; i < rank; ++i) MMR_REQUIRE(box[i] >= 1 && box[i] <= 256, "%s: box dim %d = %u", what, i, box[i]);
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s, rank %d) -> CUresult %d", what, rank, (int)r);
  return 0;
}

struct HaloPlan {
  int epi_kind = kEpiOther;
  HaloParams prm;
  void* dev_blob = nullptr;
  size_t smem_bytes = 0;
  int grid = 0;
};

}  // namespace mmr

using namespace mmr;

extern "C" int mmr_halo_conv_plan_create(const MmrHaloConvDesc* d, void** out_plan) {
  MMR_REQUIRE(d && out_plan, "null argument");
  MMR_REQUIRE(d->nsrc >= 1 && d->nsrc <= 6, "nsrc must be 1..6, got %d", d->nsrc);
  MMR_REQUIRE(d->cb == 64 || d->cb == 32 || d->cb == 16, "cb must be 16/32/64, got %d", d->cb);
  MMR_REQUIRE(d->bn >= 16 && d->bn <= 256 && d->bn % 16 == 0, "bn must be a multiple of 16 in [16,256], got %d",
              d->bn);
  MMR_REQUIRE(d->tx == 1 || d->tx == 2 || d->tx == 4, "tx must be 1, 2 or 4, got %d", d->tx);
  MMR_REQUIRE(d->tps == 1 || d->tps == 3 || d->tps == 9, "tps must be 1, 3 or 9, got %d", d->tps);
  MMR_REQUIRE(d->tps * d->bn <= 256, "tps*bn must be <= 256 (TMA box rows)");
  MMR_REQUIRE(d->acc_bufs == 1 || d->acc_bufs == 2, "acc_bufs must be 1 or 2");
  const int R = d->rph < 1 ? 1 : d->rph;
  MMR_REQUIRE(R == 1 || R == 2 || R == 4, "rph must be 1, 2 or 4, got %d", d->rph);
  MMR_REQUIRE(d->acc_bufs * d->tx * R * d->bn <= 512, "accumulators exceed 512 TMEM columns");
  if (R > 1) {
    MMR_REQUIRE(d->tps == 3 && 3 * d->bn <= 256, "row-phase stacking needs tps = 3 and bn <= 85");
    MMR_REQUIRE(d->H % R == 0, "row-phase stacking needs H %% rph == 0");
    MMR_REQUIRE(!(R == 4 && d->tx == 4), "rph = 4 with tx = 4 is not instantiated");
    MMR_REQUIRE(d->out_mode == MMR_OUT_BF16_NHWC && (d->direct_store != 0) == (d->sg < 64),
                "row-phase stacking needs the bf16 NHWC output with the default store mode of its group width");
    const bool plain = !d->scale && !d->bias && !d->residual && !d->relu;
    MMR_REQUIRE(d->stats == nullptr || (d->n_ntiles == 1 && d->bn == d->sg && plain),
                "row-phase stacking with statistics needs one channel set per CTA and a plain epilogue");
  }
  MMR_REQUIRE(d->halo_stages >= 1 && d->halo_stages <= 4, "halo_stages must be 1..4");
  MMR_REQUIRE(d->w_slots >= 1 && d->w_slots <= 8, "w_slots must be 1..8");
  MMR_REQUIRE(d->out_stages >= 1 && d->out_stages <= 2, "out_stages must be 1 or 2");
  MMR_REQUIRE(d->n_ntiles >= 1, "need at least one N tile");

  HaloPlan* pl = new HaloPlan();
  HaloParams& p = pl->prm;
  memset(&p, 0, sizeof(p));
  const int rb = d->cb * 2;
  p.rowbytes = rb;
  p.KS = rb / 32;
  p.H = d->H;
  p.W = d->W;
  p.N = d->N;
  p.TX = d->tx;
  p.R = R;
  p.bn = d->bn;
  MMR_REQUIRE((d->sg == 16 || d->sg == 32 || d->sg == 64) && d->bn % d->sg == 0,
              "sg must be 16/32/64 and divide bn (sg %d, bn %d)", d->sg, d->bn);
  p.sg = d->sg;
  p.gpn = d->bn / p.sg;
  p.tps = d->tps;
  p.nslots = 9 / d->tps;
  p.n_ntiles = d->n_ntiles;
  p.tiles_x = (d->W + 8 * d->tx - 1) / (8 * d->tx);
  p.tiles_y = (d->H + 16 * R - 1) / (16 * R);
  p.total_items = p.tiles_x * p.tiles_y * d->N * d->n_ntiles;
  p.halo_stages = d->halo_stages;
  p.w_slots = d->w_slots;
  p.acc_bufs = d->acc_bufs;
  p.out_stages = d->out_stages;
  p.pitch[0] = 8 * d->tx + 2;
  p.pitch[1] = 8 * d->tx + 4;

  std::vector<CUtensorMap> maps;
  int nchunks = 0;
  bool any_up = false;
  for (int si = 0; si < d->nsrc; ++si) {
    const MmrHaloSrc& s = d->src[si];
    MMR_REQUIRE(s.C % d->cb == 0, "source %d: %d channels are not a multiple of the chunk %d", si, s.C, d->cb);
    MMR_REQUIRE(s.N == d->N, "source %d: batch mismatch", si);
    int mi, me = -1;
    if (s.up == 1) {
      MMR_REQUIRE(s.H == d->H && s.W == d->W, "source %d: resolution mismatch", si);
      cuuint64_t dims[4] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.N};
      cuuint64_t str[3] = {(cuuint64_t)s.C * 2, (cuuint64_t)s.C * 2 * s.W, (cuuint64_t)s.C * 2 * s.W * s.H};
      cuuint32_t box[4] = {(cuuint32_t)d->cb, (cuuint32_t)p.pitch[0], (cuuint32_t)(16 * R + 2), 1};
      maps.emplace_back();
      mi = (int)maps.size() - 1;
      if (encode_generic(&maps[mi], s.ptr, 4, dims, str, box, rb, "halo source")) { delete pl; return -1; }
    } else {
      MMR_REQUIRE(s.up == 2, "source %d: up must be 1 or 2", si);
      MMR_REQUIRE(s.H * 2 == d->H && s.W * 2 == d->W, "source %d: upsampled resolution mismatch", si);
      MMR_REQUIRE(d->H % (16 * R) == 0, "nearest-x2 sources need H %% (16 * rph) == 0 (got %d)", d->H);
      any_up = true;
      const cuuint32_t bw = (cuuint32_t)(4 * d->tx + 2);
      {  // 16 interior rows: (C, dupx, Wl, dupy, N*Hl), replication through zero strides
        cuuint64_t dims[5] = {(cuuint64_t)s.C, 2, (cuuint64_t)s.W, 2, (cuuint64_t)s.N * s.H};
        cuuint64_t str[4] = {0, (cuuint64_t)s.C * 2, 0, (cuuint64_t)s.C * 2 * s.W};
        cuuint32_t box[5] = {(cuuint32_t)d->cb, 2, bw, 2, (cuuint32_t)(8 * R)};
        maps.emplace_back();
        mi = (int)maps.size() - 1;
        if (encode_generic(&maps[mi], s.ptr, 5, dims, str, box, rb, "upsampled halo body")) { delete pl; return -1; }
      }
      {  // top / bottom halo row: (C, dupx, Wl, Hl, N) keeps the per-image zero fill
        cuuint64_t dims[5] = {(cuuint64_t)s.C, 2, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.N};
        cuuint64_t str[4] = {0, (cuuint64_t)s.C * 2, (cuuint64_t)s.C * 2 * s.W, (cuuint64_t)s.C * 2 * s.W * s.H};
        cuuint32_t box[5] = {(cuuint32_t)d->cb, 2, bw, 1, 1};
        maps.emplace_back();
        me = (int)maps.size() - 1;
        if (encode_generic(&maps[me], s.ptr, 5, dims, str, box, rb, "upsampled halo edge")) { delete pl; return -1; }
      }
    }
    for (int c0 = 0; c0 < s.C; c0 += d->cb) {
      MMR_REQUIRE(nchunks < kMaxChunks, "more than %d channel chunks", kMaxChunks);
      p.chunk[nchunks++] = HaloChunk{mi, me, c0, s.up == 2 ? 1 : 0,
                                     reinterpret_cast<const uint8_t*>(s.ptr) + (size_t)c0 * 2, s.C * 2, s.W, s.H};
    }
  }
  p.nchunks = nchunks;
  {
    const int64_t rows = (int64_t)d->n_ntiles * nchunks * 9 * d->bn;
    cuuint64_t dims[2] = {(cuuint64_t)d->cb, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)rb};
    cuuint32_t box[2] = {(cuuint32_t)d->cb, (cuuint32_t)(d->tps * d->bn)};
    maps.emplace_back();
    p.wmap = (int)maps.size() - 1;
    if (encode_generic(&maps[p.wmap], d->weights, 2, dims, str, box, rb, "packed weights")) { delete pl; return -1; }
  }
  p.smap0 = (int)maps.size();
  p.out_mode = d->out_mode;
  p.direct = d->direct_store != 0;
  if (d->out_mode == MMR_OUT_BF16_NHWC) {
    const int ng = d->n_ntiles * p.gpn;
    MMR_REQUIRE(d->ngroups == ng && d->groups, "expected %d store groups, got %d", ng, d->ngroups);
    MMR_REQUIRE(ng <= kMaxGroups, "more than %d store groups", kMaxGroups);
    for (int g = 0; g < ng; ++g) {
      const MmrOutSeg& og = d->groups[g];
      MMR_REQUIRE(og.ldc % 8 == 0 && og.coff % 8 == 0 && og.coff + p.sg <= og.ldc,
                  "store group %d: channels [%d, %d) do not fit a tensor of %d channels", g, og.coff,
                  og.coff + p.sg, og.ldc);
      maps.emplace_back();
      p.group_pool[g] = og.step == -2;
      if (og.step == -2) {
        // 2x2 sum-pooled group: stored from registers, no tensor map (the entry above stays unused)
        MMR_REQUIRE(d->H % 2 == 0 && d->W % 2 == 0 && !d->stats && !d->residual && !d->bn_bwd && !d->scale && !d->bias &&
                        !d->relu,
                    "pooled store groups need even H and W and a plain epilogue");
        p.group_coff[g] = og.coff;
        p.group_ldc[g] = og.ldc;
        p.group_ptr[g] = reinterpret_cast<__nv_bfloat16*>(og.ptr);
        continue;
      }
      const int step = og.step > 1 ? og.step : 1;
      MMR_REQUIRE(step == 1 || (R == 1 && !d->direct_store && og.oy >= 0 && og.oy < step && og.ox >= 0 && og.ox < step &&
                                !d->residual && !d->stats),
                  "strided store groups need rph = 1, TMA stores, no residual / statistics and 0 <= oy, ox < step");
      if (R == 1) {
        // dense (step 1), or pixel (y, x) -> (step*y + oy, step*x + ox) of [N][step*H][step*W][ldc]
        const cuuint64_t pxb = (cuuint64_t)og.ldc * 2, rowb = pxb * d->W * step;
        cuuint64_t dims[4] = {(cuuint64_t)og.ldc, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
        cuuint64_t str[3] = {pxb * step, rowb * step, rowb * step * d->H};
        cuuint32_t box[4] = {(cuuint32_t)p.sg, 8, 4, 1};
        const uint8_t* base = reinterpret_cast<const uint8_t*>(og.ptr) + (size_t)og.oy * rowb + (size_t)og.ox * pxb;
        if (encode_generic(&maps.back(), base, 4, dims, str, box, p.sg * 2, "store group")) { delete pl; return -1; }
      } else {  // (C, W, phase, H / R, N): a warp's four tile rows are R image rows apart
        const cuuint64_t rowb = (cuuint64_t)og.ldc * 2 * d->W;
        cuuint64_t dims[5] = {(cuuint64_t)og.ldc, (cuuint64_t)d->W, (cuuint64_t)R, (cuuint64_t)(d->H / R),
                              (cuuint64_t)d->N};
        cuuint64_t str[4] = {(cuuint64_t)og.ldc * 2, rowb, rowb * R, rowb * d->H};
        cuuint32_t box[5] = {(cuuint32_t)p.sg, 8, 1, 4, 1};
        if (encode_generic(&maps.back(), og.ptr, 5, dims, str, box, p.sg * 2, "store group")) { delete pl; return -1; }
      }
      p.group_coff[g] = og.coff;
      p.group_ldc[g] = og.ldc;
      p.group_ptr[g] = reinterpret_cast<__nv_bfloat16*>(og.ptr);
    }
  } else {
    const bool metric = d->head_metric && (d->head_metric->pred_out || d->head_metric->confusion);
    MMR_REQUIRE((d->out_f32 != nullptr || metric) && d->n_ntiles == 1,
                "fp32 NCHW output needs out_f32 (or a head metric) and one N tile");
    MMR_REQUIRE(d->stats == nullptr, "statistics need the bf16 NHWC output mode");
    if (d->head_metric) {
      MMR_REQUIRE(d->bn == 16 && R == 1 && !d->scale && !d->residual && !d->relu && d->cout_total <= 16,
                  "head metric: needs the head epilogue (bn 16, rph 1, bias only, <= 16 classes)");
      MMR_REQUIRE(!d->head_metric->confusion || d->head_metric->labels, "head metric: confusion needs labels");
    }
  }
  p.cpl = d->loader == 1;
  MMR_REQUIRE(d->loader == 0 || d->loader == 1, "loader must be 0 (TMA) or 1 (cp.async)");
  MMR_REQUIRE(!p.cpl || (rb <= 64 && d->halo_stages >= 2), "the cp.async halo loader needs cb <= 32 and >= 2 halo stages");
  p.halo_tx_bytes[0] = (uint32_t)((16 * R + 2) * p.pitch[0] * rb);
  p.halo_tx_bytes[1] = (uint32_t)((16 * R + 2) * p.pitch[1] * rb);
  const uint32_t halo_bytes = p.halo_tx_bytes[any_up ? 1 : 0];
  p.halo_stage_bytes = (halo_bytes + 1023) / 1024 * 1024;
  p.w_tx_bytes = (uint32_t)(d->tps * d->bn * rb);
  p.w_slot_bytes = (p.w_tx_bytes + 1023) / 1024 * 1024;
  p.out_stage_bytes = d->out_mode == MMR_OUT_BF16_NHWC ? (uint32_t)((128 * p.sg * 2 + 1023) / 1024 * 1024) : 0;
  if (d->out_mode != MMR_OUT_BF16_NHWC) p.out_stages = 0;
  uint32_t cols = 32;
  while (cols < (uint32_t)(d->acc_bufs * d->tx * R * d->bn)) cols <<= 1;
  p.tmem_cols = cols;
  size_t smem = (size_t)p.halo_stages * p.halo_stage_bytes + (size_t)p.w_slots * p.w_slot_bytes +
                (size_t)p.out_stages * p.out_stage_bytes + 1024;  // barriers (256 B) + BN-backward mask affine (512 B)
  // head confusion counters (classes^2 x 4 B) sit after the barriers: 768 B are free there (up to 13 classes)
  if (d->out_mode == MMR_OUT_F32_NCHW && d->head_metric && d->cout_total * d->cout_total * 4 > 768) smem += 1024;
  MMR_REQUIRE(smem <= 227 * 1024, "shared memory plan needs %zu bytes (> 227 KB)", smem);
  if (smem < 116 * 1024) smem = 116 * 1024;  // one CTA per SM: the TMEM allocation assumes it
  pl->smem_bytes = smem;
  p.scale = d->scale;
  p.bias = d->bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  p.res_ldc = d->res_ldc;
  p.relu = d->relu;
  p.out_f32 = reinterpret_cast<float*>(d->out_f32);
  p.out_ldc = d->out_ldc;
  p.cout_total = d->cout_total;
  p.stats = d->stats;
  p.stats_ld = d->stats_ld;
  if (d->head_metric && d->out_mode == MMR_OUT_F32_NCHW) p.hm = *d->head_metric;
  if (const char* dbg = getenv("MMR_HALO_DBG")) p.dbg = atoi(dbg);
  if (d->bn_bwd) {
    const MmrBnBwdFused& b = *d->bn_bwd;
    const bool plain = !d->scale && !d->bias && !d->residual && !d->relu;
    MMR_REQUIRE(b.z && b.mask_scale && b.mask_shift && b.mean && b.invstd && b.coef && b.slots && b.ticket,
                "bn_bwd: z, mask_scale, mask_shift, mean, invstd, coef, slots and ticket are required");
    MMR_REQUIRE(d->out_mode == MMR_OUT_BF16_NHWC && d->n_ntiles == 1 && d->bn == d->sg && plain && !d->stats &&
                    (d->direct_store != 0) == (d->sg < 64) && d->ngroups == 1 && d->groups[0].ldc == d->sg &&
                    d->groups[0].coff == 0 && d->groups[0].step <= 1 && d->cout_total == d->sg,
                "bn_bwd needs one plain bf16 store group that is the whole destination tensor");
    p.bb = b;
    p.stats_ld = d->cout_total;
  }
  if (d->bn_finalize) {
    MMR_REQUIRE(d->stats != nullptr && d->bn_finalize->ticket != nullptr, "bn_finalize needs stats and a ticket");
    p.bnf = *d->bn_finalize;
  }

  const size_t maps_bytes = maps.size() * sizeof(CUtensorMap);
  cudaError_t e = cudaMalloc(&pl->dev_blob, maps_bytes);
  if (e != cudaSuccess) {
    delete pl;
    return fail("cudaMalloc(%zu) failed: %s", maps_bytes, cudaGetErrorString(e));
  }
  e = cudaMemcpy(pl->dev_blob, maps.data(), maps_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(pl->dev_blob);
    delete pl;
    return fail("cudaMemcpy failed: %s", cudaGetErrorString(e));
  }
  p.maps = reinterpret_cast<const CUtensorMap*>(pl->dev_blob);
  const int sms = num_sms();
  pl->grid = p.total_items < sms ? p.total_items : sms;
  {
    // the same predicates the device code used to evaluate: which epilogue family this launch runs
    const bool plain = !p.scale && !p.bias && !p.residual && !p.relu;
    const bool fast = p.out_mode == MMR_OUT_BF16_NHWC && (p.direct != 0) == (p.sg < 64) &&
                      (p.stats == nullptr || (p.n_ntiles == 1 && p.gpn == 1 && plain));
    const bool head = p.out_mode == MMR_OUT_F32_NCHW && p.bn == 16 && p.n_ntiles == 1 && p.R == 1 && !p.scale &&
                      !p.residual && !p.relu;
    if (fast && p.bb.z)
      pl->epi_kind = kEpiBnBwd;
    else if (fast && p.stats)
      pl->epi_kind = kEpiStats;
    else if (fast && plain)
      pl->epi_kind = kEpiPlain;
    else if (fast)
      pl->epi_kind = kEpiAffine;
    else if (head)
      pl->epi_kind = kEpiHead;
    else
      pl->epi_kind = kEpiOther;
    for (int g = 0; g < kMaxGroups; ++g)
      if (p.group_pool[g] && pl->epi_kind != kEpiPlain) {
        delete pl;
        return fail("pooled store groups need the plain bf16 epilogue (default store mode of the group width)");
      }
    // eight epilogue warps where the epilogue carries per-channel sums of a 64-channel store group
    // (and the scale / bias / ReLU epilogue of eval-mode and biased convs: MMR_AFFINE_EPI_WARPS=4 is the A/B switch)
    static const bool affine8 = !(getenv("MMR_AFFINE_EPI_WARPS") && atoi(getenv("MMR_AFFINE_EPI_WARPS")) == 4);
    p.epi_warps = ((pl->epi_kind == kEpiStats || pl->epi_kind == kEpiBnBwd || (pl->epi_kind == kEpiAffine && affine8)) &&
                   p.sg == 64) ? 8 : 4;
  }
  {
    const void* fns[kEpiKinds] = {(const void*)conv_halo_kernel<kEpiOther>, (const void*)conv_halo_kernel<kEpiStats>,
                                  (const void*)conv_halo_kernel<kEpiPlain>, (const void*)conv_halo_kernel<kEpiBnBwd>,
                                  (const void*)conv_halo_kernel<kEpiAffine>, (const void*)conv_halo_kernel<kEpiHead>};
    for (int k = 0; k < kEpiKinds; ++k) {
      e = cudaFuncSetAttribute(fns[k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
      if (e != cudaSuccess) cudaGetLastError();  // tolerated on non-sm_100 devices; the launch will report
    }
  }
  *out_plan = pl;
  return 0;
}

extern "C" int mmr_halo_conv_plan_run(void* plan, mmr_stream_t stream) {
  MMR_REQUIRE(plan, "null plan");
  HaloPlan* pl = reinterpret_cast<HaloPlan*>(plan);
  if (pl->prm.total_items == 0) return 0;
  switch (pl->epi_kind) {
    case kEpiStats: mmr_launch((conv_halo_kernel<kEpiStats>), pl->grid, halo_threads(kEpiStats), pl->smem_bytes, as_stream(stream), pl->prm); break;
    case kEpiPlain: mmr_launch((conv_halo_kernel<kEpiPlain>), pl->grid, kHaloThreads, pl->smem_bytes, as_stream(stream), pl->prm); break;
    case kEpiBnBwd: mmr_launch((conv_halo_kernel<kEpiBnBwd>), pl->grid, halo_threads(kEpiBnBwd), pl->smem_bytes, as_stream(stream), pl->prm); break;
    case kEpiAffine: mmr_launch((conv_halo_kernel<kEpiAffine>), pl->grid, halo_threads(kEpiAffine), pl->smem_bytes, as_stream(stream), pl->prm); break;
    case kEpiHead: mmr_launch((conv_halo_kernel<kEpiHead>), pl->grid, kHaloThreads, pl->smem_bytes, as_stream(stream), pl->prm); break;
    default: mmr_launch((conv_halo_kernel<kEpiOther>), pl->grid, kHaloThreads, pl->smem_bytes, as_stream(stream), pl->prm); break;
  }
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_debug_halo_trace(unsigned long long* out_host, int n_ctas) {
  MMR_REQUIRE(out_host && n_ctas > 0 && n_ctas <= 256, "bad argument");
  MMR_CUDA_CHECK(cudaDeviceSynchronize());
  MMR_CUDA_CHECK(cudaMemcpyFromSymbol(out_host, mmr::g_halo_trace, sizeof(unsigned long long) * n_ctas * mmr::kTraceSlots));
  return 0;
}

extern "C" int mmr_halo_conv_plan_destroy(void* plan) {
  if (!plan) return 0;
  HaloPlan* pl = reinterpret_cast<HaloPlan*>(plan);
  if (pl->dev_blob) cudaFree(pl->dev_blob);
  delete pl;
  return 0;
}

extern "C" int mmr_pack_weights_halo(const float* w_oihw, int O, int I, int mode, int cb, int bn, int n_ntiles,
                                     int nchunks, int layout, void* out, mmr_stream_t stream) {
  MMR_REQUIRE(w_oihw && out, "null argument");
  MMR_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (fprop) or 1 (dgrad)");
  const int64_t total = (int64_t)n_ntiles * nchunks * 9 * bn * cb;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  mmr_launch((pack_weights_halo_kernel), (int)blocks, 256, 0, as_stream(stream), w_oihw, O, I, mode, cb, bn, n_ntiles, nchunks, layout, reinterpret_cast<__nv_bfloat16*>(out));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_pack_items_per_block(void) { return mmr::kPackPerBlock; }

extern "C" int mmr_pack_weights_halo_batch(const MmrPackJob* jobs_dev, int njobs, int64_t total_blocks,
                                           mmr_stream_t stream) {
  MMR_REQUIRE(jobs_dev && njobs > 0 && njobs <= kMaxPackJobs && total_blocks > 0 && total_blocks < ((int64_t)1 << 31),
              "bad argument (at most %d jobs)", kMaxPackJobs);
  mmr_launch((pack_weights_halo_batch_kernel), (unsigned)total_blocks, kPackThreads, 0, as_stream(stream), jobs_dev, njobs);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}
