// Error handling, device probing and CUtensorMap encoding for libmmrseg.so.
#include <stdlib.h>

#include "common.h"

#include <cudaTypedefs.h>
#include <stdarg.h>

#include <mutex>

namespace mmr {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return -1;
}

bool pdl_enabled() {
  static const bool on = getenv("MMR_NO_PDL") == nullptr;
  return on;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    n = prop.multiProcessorCount;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static CUtensorMapSwizzle swizzle_for_bytes(int bytes) {
  if (bytes == 128) return CU_TENSOR_MAP_SWIZZLE_128B;
  if (bytes == 64) return CU_TENSOR_MAP_SWIZZLE_64B;
  return CU_TENSOR_MAP_SWIZZLE_32B;
}

void* get_encode_tiled() { return reinterpret_cast<void*>(get_encode()); }

int encode_act_map(CUtensorMap* out, const MmrSrc& s, int box_c, int box_w, int box_h, int box_n) {
  EncodeTiledFn enc = get_encode();
  MMR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  MMR_REQUIRE(box_c == 64 || box_c == 32 || box_c == 16, "activation box channels must be 16/32/64, got %d",
              box_c);
  MMR_REQUIRE(s.C % 8 == 0, "NHWC channel count must be a multiple of 8 (16-byte rows), got %d", s.C);
  MMR_REQUIRE(s.es == 1 || s.es == 2, "element stride must be 1 or 2");
  MMR_REQUIRE((reinterpret_cast<uintptr_t>(s.ptr) & 15) == 0, "tensor base must be 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.N};
  cuuint64_t strides[3] = {(cuuint64_t)s.C * 2, (cuuint64_t)s.C * 2 * s.W,
                           (cuuint64_t)s.C * 2 * s.W * s.H};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * s.es), (cuuint32_t)(box_h * s.es),
                       (cuuint32_t)box_n};
  cuuint32_t estr[4] = {1, (cuuint32_t)s.es, (cuuint32_t)s.es, 1};
  MMR_REQUIRE(box[1] <= 256 && box[2] <= 256 && box[3] <= 256, "TMA box dimension exceeds 256");
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(s.ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_c * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMR_REQUIRE(r == CUDA_SUCCESS,
              "cuTensorMapEncodeTiled(act C=%d W=%d H=%d N=%d es=%d box=%d,%d,%d,%d) -> CUresult %d",
              s.C, s.W, s.H, s.N, s.es, box_c, box_w, box_h, box_n, (int)r);
  return 0;
}

int encode_mat_map(CUtensorMap* out, const void* ptr, int rows, int cols, int box_cols,
                   int box_rows) {
  EncodeTiledFn enc = get_encode();
  MMR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  MMR_REQUIRE(cols % 8 == 0, "matrix row length must be a multiple of 8 bf16, got %d", cols);
  MMR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "matrix base must be 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_cols * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(mat %dx%d box %dx%d) -> CUresult %d", rows,
              cols, box_rows, box_cols, (int)r);
  return 0;
}

}  // namespace mmr

extern "C" {

const char* mmr_last_error(void) { return mmr::g_last_error.c_str(); }

int mmr_abi_version(void) { return 1; }

int mmr_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

}  // extern "C"
