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

}  // namespace mmr
