// Host-side helpers shared by the C-ABI translation units: error reporting and
// CUtensorMap construction through the driver entry point (no link-time libcuda
// dependency, so the library still loads on a box without a GPU driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mmrseg.h"

namespace mmr {

void set_error(const std::string& msg);
int fail(const char* fmt, ...);

#define MMR_CUDA_CHECK(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess)                                                            \
      return ::mmr::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                   \
  } while (0)

#define MMR_REQUIRE(cond, ...)                \
  do {                                        \
    if (!(cond)) return ::mmr::fail(__VA_ARGS__); \
  } while (0)

// 4-D bf16 NHWC activation map: dims (C, W, H, N), box (box_c, box_w*es, box_h*es, box_n),
// element strides (1, es, es, 1); swizzle chosen from box_c*2 bytes (128/64/32).
int encode_act_map(CUtensorMap* out, const MmrSrc& s, int box_c, int box_w, int box_h, int box_n);
// 2-D bf16 row-major matrix [rows][cols] (cols contiguous), box (box_cols, box_rows).
int encode_mat_map(CUtensorMap* out, const void* ptr, int rows, int cols, int box_cols,
                   int box_rows);

inline cudaStream_t as_stream(mmr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
int num_sms();

// Programmatic dependent launch: every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and starts with pdl_prologue().  The next
// kernel of the stream (or graph) is then scheduled while this one drains -- its launch latency, block
// scheduling and whatever precedes its own griddepcontrol.wait overlap our tail -- and blocks in
// griddepcontrol.wait until this grid has completed and its writes are visible.  wait comes before
// launch_dependents, so a kernel never runs ahead of its grandparent either.  MMR_NO_PDL=1 turns the
// attribute off (the prologue is then a no-op).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifndef MMR_PDL_EARLY_TRIGGER
#define MMR_PDL_EARLY_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_prologue() {
  pdl_wait();
#if MMR_PDL_EARLY_TRIGGER
  pdl_trigger();
#endif
}
// Late trigger: called by a CTA once its main loop is over (the MMA issuer after its last tcgen05.mma, a
// streaming kernel after its grid-stride loop).  When every CTA of the grid has got there the next kernel of
// the stream is launched: its block scheduling, barrier / TMEM set-up (everything it does before its own
// griddepcontrol.wait) overlap this grid's epilogues and block reductions.  The dependent still blocks in
// griddepcontrol.wait until this grid has completed, so the placement is a scheduling hint only.
#ifndef MMR_PDL_LATE
#define MMR_PDL_LATE 0
#endif
__device__ __forceinline__ void pdl_done() {
#if MMR_PDL_LATE
  pdl_trigger();
#endif
}
// Conv kernels: set-up that touches no global memory (mbarrier init, TMEM allocation) runs BEFORE the wait.
__device__ __forceinline__ void pdl_prologue_conv() {
#if !MMR_PDL_LATE
  pdl_prologue();
#endif
}
__device__ __forceinline__ void pdl_setup_done() {
#if MMR_PDL_LATE
  pdl_wait();
#endif
}
template <typename... KArgs, typename... Args>
inline cudaError_t mmr_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace mmr
