// Weight gradient of a convolution on tcgen05 (sm_100a).
//
//   dWt[(chunk, ch)][co] = sum_pixels x[chunk.src][pixel + shift][chunk.c0 + ch] * dz[pixel][co]
//
// Both operands are NHWC, i.e. the reduction index (pixels) is the slow one: they enter
// tcgen05.mma as MN-major operands straight from TMA boxes, no transpose anywhere.  A K-step
// is a box of 32 pixels.  The M side of one MMA stacks 128/chunk_ch activation chunks (a
// chunk is a source tensor, a channel offset and a filter tap, i.e. a shifted box); the N
// side is the conv's output channels.  One CTA holds up to 512 TMEM columns of fp32
// accumulators (several M-tiles), so every dz box it fetches is multiplied against all of
// them.  Split-K over the pixel tiles; fp32 partials are summed in a fixed order afterwards.
//
// Warp roles as in conv_gemm.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4..7 epilogue.
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace mmr {

constexpr int kWgThreads = 256;
constexpr int kWgMaxStages = 8;
constexpr int kWgPix = 32;  // pixels per K-step

struct WgParams {
  const CUtensorMap* maps;  // device: [0] dz, [1..nsrc] activation sources
  const MmrWgChunk* chunks; // device
  const MmrSrc* srcs;       // device (scalar kernel)
  MmrSrc dz;
  int dz_a, dz_bx[4], dz_by[4];
  int ncls, nchunks, chunk_ch, cout;
  int kp_w, kp_h, kp_n;
  int tiles_x, tiles_y, tiles_b, ksteps_total;
  int gx_count, gy_count, n_img;
  int n_split;
  int cpm;           // chunks per M-tile
  int n_mtiles, mt_per_group, n_groups;
  int nt_cols, n_ntiles, zb, z_boxes;
  float* partial;
  int stages;
  uint32_t z_tile_bytes, x_tile_bytes, stage_bytes, tmem_cols;
  // reduce
  float* dst;
  int dst_cout, dst_cin, dst_taps, chunk_valid_ch;
};

struct WgStep {
  int cls, n0, gy0, gx0;
};

__device__ __forceinline__ WgStep wg_decode(const WgParams& p, int it) {
  WgStep s;
  const int tx = it % p.tiles_x;
  it /= p.tiles_x;
  const int ty = it % p.tiles_y;
  it /= p.tiles_y;
  const int tb = it % p.tiles_b;
  s.cls = it / p.tiles_b;
  s.gx0 = tx * p.kp_w;
  s.gy0 = ty * p.kp_h;
  s.n0 = tb * p.kp_n;
  return s;
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  pdl_prologue_conv();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgMaxStages;
  uint64_t* tmem_full = bars + 2 * kWgMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kWgMaxStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.x / p.n_ntiles;
  const int ntile = blockIdx.x % p.n_ntiles;
  const int split = blockIdx.y;
  const int mt0 = group * p.mt_per_group;
  const int mt_cnt = min(p.mt_per_group, p.n_mtiles - mt0);
  const int k_begin = (int)(((long long)p.ksteps_total * split) / p.n_split);
  const int k_end = (int)(((long long)p.ksteps_total * (split + 1)) / p.n_split);
  const int nch = mt_cnt * p.cpm;  // activation chunks this CTA multiplies

  if (threadIdx.x == 0) {
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
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_setup_done();

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      const uint32_t bytes = p.z_boxes * p.z_tile_bytes + (uint32_t)nch * p.x_tile_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = k_begin; it < k_end; ++it) {
        const WgStep s = wg_decode(p, it);
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sz = smem + (size_t)stage * p.stage_bytes;
        uint8_t* sx = sz + p.z_boxes * p.z_tile_bytes;
        mbar_arrive_expect_tx(&full[stage], bytes);
        for (int b = 0; b < p.z_boxes; ++b)
          tma_load_4d(sz + (size_t)b * p.z_tile_bytes, &p.maps[0], &full[stage],
                      ntile * p.nt_cols + b * p.zb, p.dz_a * s.gx0 + p.dz_bx[s.cls],
                      p.dz_a * s.gy0 + p.dz_by[s.cls], s.n0);
        for (int j = 0; j < nch; ++j) {
          const MmrWgChunk* c = &p.chunks[mt0 * p.cpm + j];
          tma_load_4d(sx + (size_t)j * p.x_tile_bytes, &p.maps[1 + c->src], &full[stage], c->c0,
                      c->a * s.gx0 + c->bx[s.cls], c->a * s.gy0 + c->by[s.cls], s.n0);
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      const uint32_t idesc = make_idesc_bf16(128, p.nt_cols, 1, 1);  // both operands MN-major
      const uint32_t xrow = p.chunk_ch * 2, zrow = p.zb * 2;
      const uint32_t xswz = swizzle_code(xrow), zswz = swizzle_code(zrow);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = k_begin; it < k_end; ++it) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t z_addr = smem_u32(smem + (size_t)stage * p.stage_bytes);
        const uint32_t x_addr = z_addr + p.z_boxes * p.z_tile_bytes;
        for (int m = 0; m < mt_cnt; ++m) {
#pragma unroll
          for (int k = 0; k < kWgPix / 16; ++k) {
            const uint64_t da = make_smem_desc(x_addr + m * p.cpm * p.x_tile_bytes + k * 16 * xrow,
                                               p.x_tile_bytes, 8 * xrow, xswz);
            const uint64_t db = make_smem_desc(z_addr + k * 16 * zrow, p.z_tile_bytes, 8 * zrow, zswz);
            umma_bf16(tmem_base + (uint32_t)(m * p.nt_cols), da, db, idesc,
                      (uint32_t)(it != k_begin || k != 0));
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (k_end > k_begin)
        umma_commit(tmem_full);
      else
        mbar_arrive(tmem_full);
    }
    pdl_done();
  } else if (warp >= 4) {
    // -------------------------------------------------------------- epilogue
    const int q = warp - 4;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const size_t rows_total = (size_t)p.n_mtiles * 128;
    for (int m = 0; m < mt_cnt; ++m) {
      const size_t row = (size_t)(mt0 + m) * 128 + q * 32 + lane;
      float* dst = p.partial + ((size_t)split * rows_total + row) * p.cout + ntile * p.nt_cols;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * p.nt_cols);
      for (int c0 = 0; c0 < p.nt_cols; c0 += 16) {
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

// Scalar reference on the same tables: one thread per (row, co); writes split 0 only.
__global__ void __launch_bounds__(256)
conv_wgrad_ref_kernel(const __grid_constant__ WgParams p) {
  pdl_prologue();
  const size_t rows_total = (size_t)p.n_mtiles * 128;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_total * p.cout) return;
  const int co = (int)(idx % p.cout);
  const size_t row = idx / p.cout;
  const MmrWgChunk c = p.chunks[row / p.chunk_ch];
  const int ch = (int)(row % p.chunk_ch);
  const MmrSrc s = p.srcs[c.src];
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(s.ptr);
  const __nv_bfloat16* zp = reinterpret_cast<const __nv_bfloat16*>(p.dz.ptr);
  float acc = 0.f;
  for (int it = 0; it < p.ksteps_total; ++it) {
    const WgStep st = wg_decode(p, it);
    for (int pix = 0; pix < kWgPix; ++pix) {
      const int w = pix % p.kp_w, h = (pix / p.kp_w) % p.kp_h, b = pix / (p.kp_w * p.kp_h);
      const int n = st.n0 + b;
      const int zx = p.dz_a * st.gx0 + p.dz_bx[st.cls] + w * p.dz.es;
      const int zy = p.dz_a * st.gy0 + p.dz_by[st.cls] + h * p.dz.es;
      const int xx = c.a * st.gx0 + c.bx[st.cls] + w * s.es;
      const int xy = c.a * st.gy0 + c.by[st.cls] + h * s.es;
      if (n >= p.dz.N || zx < 0 || zx >= p.dz.W || zy < 0 || zy >= p.dz.H) continue;
      if (n >= s.N || xx < 0 || xx >= s.W || xy < 0 || xy >= s.H) continue;
      if (c.c0 + ch >= s.C || co >= p.dz.C) continue;
      acc += __bfloat162float(xp[(((size_t)n * s.H + xy) * s.W + xx) * s.C + c.c0 + ch]) *
             __bfloat162float(zp[(((size_t)n * p.dz.H + zy) * p.dz.W + zx) * p.dz.C + co]);
    }
  }
  p.partial[row * p.cout + co] = acc;
}

// Sum the split-K partials in split order and scatter into OIHW fp32.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const __grid_constant__ WgParams p, int n_split, int accumulate) {
  pdl_prologue();
  const size_t rows_total = (size_t)p.n_mtiles * 128;
  const size_t total = rows_total * p.cout;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(idx % p.cout);
    const size_t row = idx / p.cout;
    const MmrWgChunk* c = &p.chunks[row / p.chunk_ch];
    const int ch = (int)(row % p.chunk_ch);
    const int tap = c->dst_tap;
    const int ci = c->dst_ci + ch;
    if (tap < 0 || ch >= p.chunk_valid_ch || ci >= p.dst_cin || co >= p.dst_cout) continue;
    float s = 0.f;
    for (int k = 0; k < n_split; ++k) s += __ldg(p.partial + (size_t)k * total + idx);
    float* d = p.dst + ((size_t)co * p.dst_cin + ci) * p.dst_taps + tap;
    *d = accumulate ? *d + s : s;
  }
}

struct WgPlan {
  WgParams prm;
  void* dev_blob = nullptr;
  size_t smem_bytes = 0;
};

static uint32_t wg_round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

}  // namespace mmr

using namespace mmr;

extern "C" int mmr_wgrad_plan_create(const MmrWgradDesc* d, void** out_plan) {
  MMR_REQUIRE(d && out_plan, "null argument");
  MMR_REQUIRE(d->nsrc >= 1 && d->nsrc <= 6, "wgrad: nsrc must be 1..6");
  MMR_REQUIRE(d->chunk_ch == 64 || d->chunk_ch == 32 || d->chunk_ch == 16,
              "wgrad: chunk_ch must be 16/32/64, got %d", d->chunk_ch);
  const int cpm = 128 / d->chunk_ch;
  MMR_REQUIRE(d->nchunks > 0 && d->nchunks % cpm == 0, "wgrad: nchunks (%d) must be a multiple of %d",
              d->nchunks, cpm);
  const int co = d->cout_gemm;
  MMR_REQUIRE(co == 16 || co == 32 || co == 64 || co == 128 || co == 256 || co == 512,
              "wgrad: cout_gemm must be 16/32/64/128/256/512, got %d", co);
  MMR_REQUIRE(d->dz.C >= co, "wgrad: dz has %d channels, GEMM needs %d", d->dz.C, co);
  MMR_REQUIRE(d->kp_w * d->kp_h * d->kp_n == kWgPix, "wgrad: kp_w*kp_h*kp_n must be 32");
  MMR_REQUIRE(d->ncls >= 1 && d->ncls <= 4, "wgrad: ncls must be 1..4");
  MMR_REQUIRE(d->partial && d->dst, "wgrad: null output");
  for (int i = 0; i < d->nchunks; ++i)
    MMR_REQUIRE(d->chunks[i].src >= 0 && d->chunks[i].src < d->nsrc, "wgrad: chunk %d bad source", i);

  WgPlan* pl = new WgPlan();
  WgParams& p = pl->prm;
  memset(&p, 0, sizeof(p));
  p.nt_cols = co > 256 ? 256 : co;
  p.n_ntiles = co / p.nt_cols;
  p.zb = p.nt_cols < 64 ? p.nt_cols : 64;
  p.z_boxes = p.nt_cols / p.zb;

  std::vector<CUtensorMap> maps(d->nsrc + 1);
  if (encode_act_map(&maps[0], d->dz, p.zb, d->kp_w, d->kp_h, d->kp_n) != 0) {
    delete pl;
    return -1;
  }
  for (int i = 0; i < d->nsrc; ++i)
    if (encode_act_map(&maps[1 + i], d->src[i], d->chunk_ch, d->kp_w, d->kp_h, d->kp_n) != 0) {
      delete pl;
      return -1;
    }
  const size_t maps_bytes = maps.size() * sizeof(CUtensorMap);
  const size_t ch_bytes = wg_round_up((uint32_t)(d->nchunks * sizeof(MmrWgChunk)), 128);
  const size_t src_bytes = wg_round_up((uint32_t)(d->nsrc * sizeof(MmrSrc)), 128);
  const size_t total = maps_bytes + ch_bytes + src_bytes;
  std::vector<uint8_t> host(total, 0);
  memcpy(host.data(), maps.data(), maps_bytes);
  memcpy(host.data() + maps_bytes, d->chunks, d->nchunks * sizeof(MmrWgChunk));
  memcpy(host.data() + maps_bytes + ch_bytes, d->src, d->nsrc * sizeof(MmrSrc));
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
  p.chunks = reinterpret_cast<const MmrWgChunk*>(blob + maps_bytes);
  p.srcs = reinterpret_cast<const MmrSrc*>(blob + maps_bytes + ch_bytes);
  p.dz = d->dz;
  p.dz_a = d->dz_a;
  for (int c = 0; c < 4; ++c) p.dz_bx[c] = d->dz_bx[c], p.dz_by[c] = d->dz_by[c];
  p.ncls = d->ncls;
  p.nchunks = d->nchunks;
  p.chunk_ch = d->chunk_ch;
  p.cout = co;
  p.kp_w = d->kp_w;
  p.kp_h = d->kp_h;
  p.kp_n = d->kp_n;
  p.gx_count = d->gx_count;
  p.gy_count = d->gy_count;
  p.n_img = d->n_img;
  p.tiles_x = (d->gx_count + d->kp_w - 1) / d->kp_w;
  p.tiles_y = (d->gy_count + d->kp_h - 1) / d->kp_h;
  p.tiles_b = (d->n_img + d->kp_n - 1) / d->kp_n;
  p.ksteps_total = p.tiles_x * p.tiles_y * p.tiles_b * d->ncls;
  p.n_split = d->n_split < 1 ? 1 : d->n_split;
  if (p.n_split > p.ksteps_total) p.n_split = p.ksteps_total;
  p.cpm = cpm;
  p.n_mtiles = d->nchunks / cpm;
  const int max_mt = 512 / p.nt_cols > 8 ? 8 : 512 / p.nt_cols;
  p.n_groups = (p.n_mtiles + max_mt - 1) / max_mt;
  p.mt_per_group = (p.n_mtiles + p.n_groups - 1) / p.n_groups;
  p.n_groups = (p.n_mtiles + p.mt_per_group - 1) / p.mt_per_group;
  p.partial = d->partial;
  p.z_tile_bytes = (uint32_t)(kWgPix * p.zb * 2);
  p.x_tile_bytes = (uint32_t)(kWgPix * d->chunk_ch * 2);
  p.stage_bytes = wg_round_up(p.z_boxes * p.z_tile_bytes + p.mt_per_group * cpm * p.x_tile_bytes, 1024);
  int stages = (int)((200 * 1024) / p.stage_bytes);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  MMR_REQUIRE(stages >= 2, "wgrad: stage of %u bytes does not fit twice in shared memory", p.stage_bytes);
  p.stages = stages;
  uint32_t cols = 32;
  while (cols < (uint32_t)(p.mt_per_group * p.nt_cols)) cols <<= 1;
  p.tmem_cols = cols;
  p.dst = d->dst;
  p.dst_cout = d->dst_cout;
  p.dst_cin = d->dst_cin;
  p.dst_taps = d->dst_taps;
  p.chunk_valid_ch = d->chunk_valid_ch > 0 ? d->chunk_valid_ch : d->chunk_ch;
  size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + 256;
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM (TMEM budget)
  pl->smem_bytes = smem;
  e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(220 * 1024));
  if (e != cudaSuccess) cudaGetLastError();
  *out_plan = pl;
  return 0;
}

extern "C" int mmr_wgrad_plan_run(void* plan, int impl, int accumulate, mmr_stream_t stream) {
  MMR_REQUIRE(plan, "null plan");
  WgPlan* pl = reinterpret_cast<WgPlan*>(plan);
  const WgParams& p = pl->prm;
  const size_t total = (size_t)p.n_mtiles * 128 * p.cout;
  int n_split = p.n_split;
  if (impl == 0) {
    dim3 grid(p.n_groups * p.n_ntiles, p.n_split);
    mmr_launch((conv_wgrad_tc_kernel), grid, kWgThreads, pl->smem_bytes, as_stream(stream), p);
  } else {
    mmr_launch((conv_wgrad_ref_kernel), (unsigned)((total + 255) / 256), 256, 0, as_stream(stream), p);
    n_split = 1;
  }
  MMR_CUDA_CHECK(cudaGetLastError());
  int64_t blocks = (int64_t)((total + 255) / 256);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  mmr_launch((wgrad_reduce_kernel), (int)blocks, 256, 0, as_stream(stream), p, n_split, accumulate);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_wgrad_plan_destroy(void* plan) {
  if (!plan) return 0;
  WgPlan* pl = reinterpret_cast<WgPlan*>(plan);
  if (pl->dev_blob) cudaFree(pl->dev_blob);
  delete pl;
  return 0;
}
