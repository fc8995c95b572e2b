// Implicit-GEMM convolution for sm_100a: TMA-gathered NHWC activation boxes and K-major
// weight blocks feed tcgen05.mma with fp32 accumulators in TMEM; a 4-warp epilogue applies
// scale/bias/residual/ReLU and stores bf16 NHWC (or fp32 NCHW logits).
//
// One persistent CTA per SM, 8 warps:
//   warp 0 (1 lane)  TMA producer: walks the K-step table of the tile's pixel class
//   warp 1 (1 lane)  MMA issuer:   bk/16 tcgen05.mma per K-step, commit -> frees the stage
//   warp 2           TMEM allocator (2 accumulators of bn columns: MMA of tile i+1
//                    overlaps the epilogue of tile i)
//   warps 4..7       epilogue: tcgen05.ld 32 lanes x 16 columns at a time
//
// The same kernel runs forward convolutions (3x3/1x1, stride 1/2, nearest-x2 + concat
// K-segments) and data gradients (dgrad = the same gather with [Cin][tap][Cout] weights);
// what differs is only the host-built K-step table.  See include/mmrseg.h.
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace mmr {

constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kTileM = 128;

struct ConvParams {
  const CUtensorMap* maps;  // device array: nsrc activation maps, then the weight map
  const MmrKStep* ksteps;   // device
  const MmrOutSeg* outsegs; // device
  const MmrSrc* srcs;       // device (scalar reference kernel only)
  const __nv_bfloat16* weights;
  int w_rows, w_cols;
  MmrConvClass cls[4];
  int ncls, nsrc;
  int bk, bn;
  int box_w, box_h, box_n;
  int tiles_x, tiles_y, tiles_b, n_tiles_n, total_tiles;
  int gx_count, gy_count, n_img;
  int oy_mul, ox_mul, Hout, Wout, cout_total;
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;
  int res_ldc, relu, out_mode;
  int stages;
  uint32_t a_bytes, b_bytes, stage_bytes;
  uint32_t tmem_cols;
};

struct TileCoord {
  int cls, tb, ty, tx, nt;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t) {
  TileCoord c;
  c.nt = t % p.n_tiles_n;
  t /= p.n_tiles_n;
  c.tx = t % p.tiles_x;
  t /= p.tiles_x;
  c.ty = t % p.tiles_y;
  t /= p.tiles_y;
  c.tb = t % p.tiles_b;
  c.cls = t / p.tiles_b;
  return c;
}

// Epilogue for 16 consecutive output channels of one pixel row.
__device__ __forceinline__ void store_row16(const ConvParams& p, const MmrOutSeg& seg, int nt, int c0,
                                            const float (&acc)[16], int n, int y, int x,
                                            int ncols_valid) {
  const int ch0 = nt * p.bn + c0;  // global output channel of acc[0]
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float s = p.scale ? __ldg(p.scale + min(ch0 + j, p.cout_total - 1)) : 1.f;
    float b = p.bias ? __ldg(p.bias + min(ch0 + j, p.cout_total - 1)) : 0.f;
    v[j] = acc[j] * s + b;
  }
  const size_t pix = ((size_t)n * p.Hout + y) * p.Wout + x;
  if (p.residual) {
    const uint4* r = reinterpret_cast<const uint4*>(p.residual + pix * p.res_ldc + ch0);
    uint4 r0 = __ldg(r), r1 = __ldg(r + 1);
    uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float2 f = unpack_bf16x2(rr[j]);
      v[2 * j] += f.x;
      v[2 * j + 1] += f.y;
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (p.out_mode == MMR_OUT_BF16_NHWC) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(seg.ptr) + pix * seg.ldc + seg.coff + c0;
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]);
    o0.y = pack_bf16x2(v[2], v[3]);
    o0.z = pack_bf16x2(v[4], v[5]);
    o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]);
    o1.y = pack_bf16x2(v[10], v[11]);
    o1.z = pack_bf16x2(v[12], v[13]);
    o1.w = pack_bf16x2(v[14], v[15]);
    reinterpret_cast<uint4*>(dst)[0] = o0;
    reinterpret_cast<uint4*>(dst)[1] = o1;
  } else {
    float* dst = reinterpret_cast<float*>(seg.ptr);
    const size_t hw = (size_t)p.Hout * p.Wout;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (c0 + j < ncols_valid) {
        const int ch = seg.coff + c0 + j;
        dst[((size_t)n * seg.ldc + ch) * hw + (size_t)y * p.Wout + x] = v[j];
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ ConvParams p) {
  pdl_prologue_conv();
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: swizzle-128B atoms and the UMMA descriptors assume it.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, p.tmem_cols);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i <= p.nsrc; ++i) tma_prefetch_desc(&p.maps[i]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_setup_done();

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        const MmrConvClass cl = p.cls[tc.cls];
        const int gx0 = tc.tx * p.box_w, gy0 = tc.ty * p.box_h, n0 = tc.tb * p.box_n;
        for (int ks = 0; ks < cl.kcount; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1);
          const MmrKStep k = p.ksteps[cl.kbegin + ks];
          uint8_t* sa = stage_base + (size_t)stage * p.stage_bytes;
          uint8_t* sb = sa + p.a_bytes;
          mbar_arrive_expect_tx(&full[stage], p.a_bytes + p.b_bytes);
          tma_load_4d(sa, &p.maps[k.src], &full[stage], k.c0, k.ax * gx0 + k.bx,
                      k.ay * gy0 + k.by, n0);
          tma_load_2d(sb, &p.maps[p.nsrc], &full[stage], k.wk, tc.nt * p.bn);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      const uint32_t idesc = make_idesc_bf16(kTileM, p.bn, 0, 0);
      const uint32_t row_bytes = p.bk * 2;
      const uint32_t sbo = 8 * row_bytes;
      const uint32_t swz = swizzle_code(row_bytes);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const TileCoord tc = decode_tile(p, t);
        const MmrConvClass cl = p.cls[tc.cls];
        const int acc = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(&tmem_empty[acc], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.bn);
        for (int ks = 0; ks < cl.kcount; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stage_base + (size_t)stage * p.stage_bytes);
          const uint32_t b_addr = a_addr + p.a_bytes;
          for (int k = 0; k < p.bk / 16; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * 32, 16, sbo, swz);
            const uint64_t db = make_smem_desc(b_addr + k * 32, 16, sbo, swz);
            umma_bf16(d_tmem, da, db, idesc, (uint32_t)((ks | k) != 0));
          }
          umma_commit(&empty[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (cl.kcount > 0)
          umma_commit(&tmem_full[acc]);
        else
          mbar_arrive(&tmem_full[acc]);
      }
    }
    pdl_done();
  } else if (warp >= 4) {
    // -------------------------------------------------------------- epilogue
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const int per_img = p.box_w * p.box_h;
    const int bi = row / per_img;
    const int rem = row - bi * per_img;
    const int h = rem / p.box_w;
    const int w = rem - h * p.box_w;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const TileCoord tc = decode_tile(p, t);
      const MmrConvClass cl = p.cls[tc.cls];
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(&tmem_full[acc], use & 1);
      tc_fence_after();
      const int gx = tc.tx * p.box_w + w, gy = tc.ty * p.box_h + h, n = tc.tb * p.box_n + bi;
      const bool valid = gx < p.gx_count && gy < p.gy_count && n < p.n_img;
      const int y = p.oy_mul * gy + cl.oy_add, x = p.ox_mul * gx + cl.ox_add;
      const MmrOutSeg seg = p.outsegs[tc.nt];
      const int ncols_valid = min(p.bn, p.cout_total - tc.nt * p.bn);
      const uint32_t taddr = tmem_base + (uint32_t)(acc * p.bn) + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        uint32_t r[16];
        if (cl.kcount > 0) {
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
        if (valid && c0 < ncols_valid) {
          float a[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) a[j] = __uint_as_float(r[j]);
          store_row16(p, seg, tc.nt, c0, a, n, y, x, ncols_valid);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// Scalar reference: one CTA per tile, one thread per pixel row, same tables and epilogue.
__global__ void __launch_bounds__(kTileM)
conv_gemm_ref_kernel(const __grid_constant__ ConvParams p) {
  pdl_prologue();
  const int t = blockIdx.x;
  if (t >= p.total_tiles) return;
  const TileCoord tc = decode_tile(p, t);
  const MmrConvClass cl = p.cls[tc.cls];
  const int row = threadIdx.x;
  const int per_img = p.box_w * p.box_h;
  const int bi = row / per_img;
  const int rem = row - bi * per_img;
  const int h = rem / p.box_w;
  const int w = rem - h * p.box_w;
  const int gx0 = tc.tx * p.box_w, gy0 = tc.ty * p.box_h, n0 = tc.tb * p.box_n;
  const int gx = gx0 + w, gy = gy0 + h, n = n0 + bi;
  const bool valid = gx < p.gx_count && gy < p.gy_count && n < p.n_img;
  if (!valid) return;
  const int y = p.oy_mul * gy + cl.oy_add, x = p.ox_mul * gx + cl.ox_add;
  const MmrOutSeg seg = p.outsegs[tc.nt];
  const int ncols_valid = min(p.bn, p.cout_total - tc.nt * p.bn);
  for (int c0 = 0; c0 < p.bn && c0 < ncols_valid; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int ks = 0; ks < cl.kcount; ++ks) {
      const MmrKStep k = p.ksteps[cl.kbegin + ks];
      const MmrSrc s = p.srcs[k.src];
      const int xs = k.ax * gx0 + k.bx + w * s.es;
      const int ys = k.ay * gy0 + k.by + h * s.es;
      const bool inb = xs >= 0 && xs < s.W && ys >= 0 && ys < s.H && n < s.N;
      if (!inb) continue;
      const __nv_bfloat16* a =
          reinterpret_cast<const __nv_bfloat16*>(s.ptr) + (((size_t)n * s.H + ys) * s.W + xs) * s.C;
      for (int kk = 0; kk < p.bk; ++kk) {
        const int c = k.c0 + kk;
        if (c >= s.C) break;
        const float av = __bfloat162float(a[c]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int wr = tc.nt * p.bn + c0 + j;
          if (wr < p.w_rows && k.wk + kk < p.w_cols)
            acc[j] += av * __bfloat162float(p.weights[(size_t)wr * p.w_cols + k.wk + kk]);
        }
      }
    }
    store_row16(p, seg, tc.nt, c0, acc, n, y, x, ncols_valid);
  }
}

struct ConvPlan {
  ConvParams prm;
  void* dev_blob = nullptr;
  size_t smem_bytes = 0;
  int grid = 0;
};

static uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

}  // namespace mmr

using namespace mmr;

extern "C" int mmr_conv_plan_create(const MmrConvDesc* d, void** out_plan) {
  MMR_REQUIRE(d && out_plan, "null argument");
  MMR_REQUIRE(d->nsrc >= 1 && d->nsrc <= 6, "nsrc must be 1..6, got %d", d->nsrc);
  MMR_REQUIRE(d->bk == 64 || d->bk == 32 || d->bk == 16, "bk must be 16/32/64, got %d", d->bk);
  MMR_REQUIRE(d->bn >= 16 && d->bn <= 256 && d->bn % 16 == 0, "bn must be a multiple of 16 in [16,256], got %d",
              d->bn);
  MMR_REQUIRE(d->box_w * d->box_h * d->box_n == kTileM, "box_w*box_h*box_n must be 128, got %d*%d*%d",
              d->box_w, d->box_h, d->box_n);
  MMR_REQUIRE(d->ncls >= 1 && d->ncls <= 4, "ncls must be 1..4");
  MMR_REQUIRE(d->n_tiles_n >= 1 && d->outsegs, "need at least one N-tile");
  MMR_REQUIRE(d->out_mode == MMR_OUT_F32_NCHW || d->cout_total % 16 == 0,
              "bf16 NHWC output needs cout_total %% 16 == 0, got %d", d->cout_total);
  MMR_REQUIRE(d->w_cols % 8 == 0, "weight matrix row length must be a multiple of 8");
  for (int i = 0; i < d->nksteps; ++i) {
    const MmrKStep& k = d->ksteps[i];
    MMR_REQUIRE(k.src >= 0 && k.src < d->nsrc, "K-step %d: bad source %d", i, k.src);
    MMR_REQUIRE(k.wk >= 0 && k.wk < d->w_cols, "K-step %d: bad weight column %d", i, k.wk);
    MMR_REQUIRE(k.c0 >= 0 && k.c0 < d->src[k.src].C, "K-step %d: bad channel offset %d", i, k.c0);
  }
  for (int c = 0; c < d->ncls; ++c)
    MMR_REQUIRE(d->cls[c].kbegin >= 0 && d->cls[c].kbegin + d->cls[c].kcount <= d->nksteps,
                "class %d K-step range out of bounds", c);

  ConvPlan* pl = new ConvPlan();
  ConvParams& p = pl->prm;
  memset(&p, 0, sizeof(p));

  std::vector<CUtensorMap> maps(d->nsrc + 1);
  for (int i = 0; i < d->nsrc; ++i) {
    if (encode_act_map(&maps[i], d->src[i], d->bk, d->box_w, d->box_h, d->box_n) != 0) {
      delete pl;
      return -1;
    }
  }
  if (encode_mat_map(&maps[d->nsrc], d->weights, d->w_rows, d->w_cols, d->bk, d->bn) != 0) {
    delete pl;
    return -1;
  }

  // One device blob: maps (128-byte aligned) | ksteps | outsegs | srcs.
  const size_t maps_bytes = maps.size() * sizeof(CUtensorMap);
  const size_t ks_bytes = round_up((uint32_t)(d->nksteps * sizeof(MmrKStep)), 128);
  const size_t seg_bytes = round_up((uint32_t)(d->n_tiles_n * sizeof(MmrOutSeg)), 128);
  const size_t src_bytes = round_up((uint32_t)(d->nsrc * sizeof(MmrSrc)), 128);
  const size_t total = maps_bytes + ks_bytes + seg_bytes + src_bytes;
  std::vector<uint8_t> host(total, 0);
  memcpy(host.data(), maps.data(), maps_bytes);
  if (d->nksteps) memcpy(host.data() + maps_bytes, d->ksteps, d->nksteps * sizeof(MmrKStep));
  memcpy(host.data() + maps_bytes + ks_bytes, d->outsegs, d->n_tiles_n * sizeof(MmrOutSeg));
  memcpy(host.data() + maps_bytes + ks_bytes + seg_bytes, d->src, d->nsrc * sizeof(MmrSrc));
  cudaError_t e = cudaMalloc(&pl->dev_blob, total);
  if (e != cudaSuccess) {
    delete pl;
    return fail("cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e));
  }
  e = cudaMemcpy(pl->dev_blob, host.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(pl->dev_blob);
    delete pl;
    return fail("cudaMemcpy failed: %s", cudaGetErrorString(e));
  }
  uint8_t* blob = reinterpret_cast<uint8_t*>(pl->dev_blob);
  p.maps = reinterpret_cast<const CUtensorMap*>(blob);
  p.ksteps = reinterpret_cast<const MmrKStep*>(blob + maps_bytes);
  p.outsegs = reinterpret_cast<const MmrOutSeg*>(blob + maps_bytes + ks_bytes);
  p.srcs = reinterpret_cast<const MmrSrc*>(blob + maps_bytes + ks_bytes + seg_bytes);
  p.weights = reinterpret_cast<const __nv_bfloat16*>(d->weights);
  p.w_rows = d->w_rows;
  p.w_cols = d->w_cols;
  for (int c = 0; c < 4; ++c) p.cls[c] = d->cls[c];
  p.ncls = d->ncls;
  p.nsrc = d->nsrc;
  p.bk = d->bk;
  p.bn = d->bn;
  p.box_w = d->box_w;
  p.box_h = d->box_h;
  p.box_n = d->box_n;
  p.tiles_x = (d->gx_count + d->box_w - 1) / d->box_w;
  p.tiles_y = (d->gy_count + d->box_h - 1) / d->box_h;
  p.tiles_b = (d->n_img + d->box_n - 1) / d->box_n;
  p.n_tiles_n = d->n_tiles_n;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_b * p.n_tiles_n * d->ncls;
  p.gx_count = d->gx_count;
  p.gy_count = d->gy_count;
  p.n_img = d->n_img;
  p.oy_mul = d->oy_mul;
  p.ox_mul = d->ox_mul;
  p.Hout = d->Hout;
  p.Wout = d->Wout;
  p.cout_total = d->cout_total;
  p.scale = d->scale;
  p.bias = d->bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  p.res_ldc = d->res_ldc;
  p.relu = d->relu;
  p.out_mode = d->out_mode;
  // Stage = A box (128 x bk bf16, always a multiple of 1024 B) + B slot (bn x bk bf16 padded to
  // 1024 B so every operand base keeps the swizzle-atom alignment).  The mbarrier expects the
  // bytes TMA really writes: the unpadded boxes.
  p.a_bytes = (uint32_t)(kTileM * d->bk * 2);
  p.b_bytes = (uint32_t)(d->bn * d->bk * 2);
  p.stage_bytes = p.a_bytes + round_up(p.b_bytes, 1024);
  const size_t budget = 200 * 1024;
  int stages = (int)(budget / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * d->bn)) cols <<= 1;
  p.tmem_cols = cols;
  size_t smem = (size_t)p.stages * p.stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM: TMEM and the persistent schedule assume it
  pl->smem_bytes = smem;
  const int sms = num_sms();
  pl->grid = p.total_tiles < sms ? p.total_tiles : sms;
  e = cudaFuncSetAttribute(conv_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(220 * 1024));
  if (e != cudaSuccess) {
    cudaGetLastError();  // tolerated on non-sm_100 devices; the launch will report
  }
  *out_plan = pl;
  return 0;
}

extern "C" int mmr_conv_plan_run(void* plan, int impl, mmr_stream_t stream) {
  MMR_REQUIRE(plan, "null plan");
  ConvPlan* pl = reinterpret_cast<ConvPlan*>(plan);
  if (pl->prm.total_tiles == 0) return 0;
  if (impl == 0) {
    mmr_launch((conv_gemm_tc_kernel), pl->grid, kThreads, pl->smem_bytes, as_stream(stream), pl->prm);
  } else {
    mmr_launch((conv_gemm_ref_kernel), pl->prm.total_tiles, kTileM, 0, as_stream(stream), pl->prm);
  }
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_conv_plan_destroy(void* plan) {
  if (!plan) return 0;
  ConvPlan* pl = reinterpret_cast<ConvPlan*>(plan);
  if (pl->dev_blob) cudaFree(pl->dev_blob);
  delete pl;
  return 0;
}
