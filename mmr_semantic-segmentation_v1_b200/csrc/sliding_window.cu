// Sliding-window inference on the device: the reference's
// monai.inferers.sliding_window_inference(frames, roi_size, sw_batch_size, predictor, overlap,
// mode="constant") call at ED/Main_MMR_SegModel.py:1308-1317 followed by `preds.argmax(1)` (:1320).
//
//   window_gather : frames [N][3][H][W] fp32 (or uint8 [N][H][W][3]) -> a batch of windows in the
//                   predictor's input layout, window w = (frame, y0, x0) from the start lists;
//   window_blend  : window logits [windows][C][rh][rw] fp32 -> out [N][C][H][W] = mean over the windows
//                   covering each pixel (constant importance map), summed in window order by ONE
//                   thread per pixel -- deterministic, no atomics -- with the argmax (first maximum,
//                   torch's rule) of the blended logits fused in, so the blended map never has to be
//                   re-read for the metric.
#include "common.h"

namespace mmr {

constexpr int kMaxStarts = 16;

struct WindowGrid {
  int ny, nx;            // window starts per axis; windows of one frame are ordered (iy, ix)
  int ys[kMaxStarts];
  int xs[kMaxStarts];
};

template <typename TIn>
__global__ void window_gather_kernel(const TIn* __restrict__ frames, int N, int H, int W, WindowGrid g, int rh,
                                     int rw, int w_begin, int w_count, TIn* __restrict__ out) {
  pdl_prologue();
  // fp32: NCHW in, NCHW out.  uint8: NHWC in, NHWC out.
  const int wpf = g.ny * g.nx;
  const int64_t per = (int64_t)3 * rh * rw;
  const int64_t total = per * w_count;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int wl = (int)(i / per);
    const int64_t r = i % per;
    const int w = w_begin + wl;
    TIn v = (TIn)0;
    if (w < N * wpf) {   // windows past the last one (padding of the final batch) are zero
      const int n = w / wpf, k = w % wpf;
      const int y0 = g.ys[k / g.nx], x0 = g.xs[k % g.nx];
      if (sizeof(TIn) == 1) {
        const int c = (int)(r % 3), x = (int)((r / 3) % rw), y = (int)(r / (3 * (int64_t)rw));
        v = frames[(((size_t)n * H + y0 + y) * W + x0 + x) * 3 + c];
      } else {
        const int x = (int)(r % rw), y = (int)((r / rw) % rh), c = (int)(r / ((int64_t)rw * rh));
        v = frames[(((size_t)n * 3 + c) * H + y0 + y) * W + x0 + x];
      }
    }
    out[i] = v;
  }
}

__global__ void window_blend_kernel(const float* __restrict__ win, int N, int C, int H, int W, WindowGrid g, int rh,
                                    int rw, float* __restrict__ out, int64_t* __restrict__ pred) {
  pdl_prologue();
  const int wpf = g.ny * g.nx;
  const int64_t hw = (int64_t)H * W;
  const int64_t total = (int64_t)N * hw;
  const size_t wplane = (size_t)rh * rw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / hw);
    const int64_t p = i % hw;
    const int y = (int)(p / W), x = (int)(p % W);
    // covering windows, in window order
    int cover[kMaxStarts * 4];
    int ncov = 0;
    for (int iy = 0; iy < g.ny; ++iy) {
      const int dy = y - g.ys[iy];
      if (dy < 0 || dy >= rh) continue;
      for (int ix = 0; ix < g.nx; ++ix) {
        const int dx = x - g.xs[ix];
        if (dx < 0 || dx >= rw) continue;
        if (ncov < kMaxStarts * 4) cover[ncov++] = ((iy * g.nx + ix) << 20) | (dy * rw + dx);
      }
    }
    const float cnt = (float)ncov;   // >= 1: the start lists cover every pixel
    float best = 0.f;
    int bi = 0;
    bool best_nan = false;
    for (int c = 0; c < C; ++c) {
      float s = 0.f;
      for (int k = 0; k < ncov; ++k) {
        const int w = n * wpf + (cover[k] >> 20);
        s += __ldg(win + ((size_t)w * C + c) * wplane + (cover[k] & 0xFFFFF));
      }
      // monai: accumulated window sum divided by the accumulated (constant) importance count
      const float v = s / cnt;
      if (out) out[((size_t)n * C + c) * hw + p] = v;
      // torch.argmax: first maximum; a NaN is maximal and the first NaN wins
      if (!best_nan && (c == 0 || v > best || v != v)) {
        best = v;
        bi = c;
        best_nan = v != v;
      }
    }
    if (pred) pred[i] = bi;
  }
}

static int fill_grid(WindowGrid& g, const int* ys, int ny, const int* xs, int nx, int H, int W, int rh, int rw) {
  MMR_REQUIRE(ny >= 1 && ny <= kMaxStarts && nx >= 1 && nx <= kMaxStarts, "1..%d window starts per axis", kMaxStarts);
  MMR_REQUIRE(rh >= 1 && rw >= 1 && rh <= H && rw <= W && (int64_t)rh * rw < (1 << 20),
              "window %dx%d must fit the %dx%d frame (and hold fewer than 2^20 pixels)", rh, rw, H, W);
  g.ny = ny;
  g.nx = nx;
  for (int i = 0; i < kMaxStarts; ++i) {
    g.ys[i] = i < ny ? ys[i] : 0;
    g.xs[i] = i < nx ? xs[i] : 0;
  }
  for (int i = 0; i < ny; ++i) MMR_REQUIRE(ys[i] >= 0 && ys[i] + rh <= H, "window row start %d out of range", ys[i]);
  for (int i = 0; i < nx; ++i) MMR_REQUIRE(xs[i] >= 0 && xs[i] + rw <= W, "window column start %d out of range", xs[i]);
  return 0;
}

static int sw_blocks(int64_t work) {
  int64_t b = (work + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 32;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace mmr

using namespace mmr;

extern "C" int mmr_window_gather(const void* frames, int is_u8, int N, int H, int W, const int* ys, int ny,
                                 const int* xs, int nx, int rh, int rw, int w_begin, int w_count, void* out,
                                 mmr_stream_t stream) {
  MMR_REQUIRE(frames && out && ys && xs && w_count >= 1 && w_begin >= 0, "bad argument");
  WindowGrid g;
  if (fill_grid(g, ys, ny, xs, nx, H, W, rh, rw)) return -1;
  const int64_t total = (int64_t)3 * rh * rw * w_count;
  if (is_u8)
    mmr_launch((window_gather_kernel<uint8_t>), sw_blocks(total), 256, 0, as_stream(stream), reinterpret_cast<const uint8_t*>(frames), N, H, W, g, rh, rw, w_begin, w_count, reinterpret_cast<uint8_t*>(out));
  else
    mmr_launch((window_gather_kernel<float>), sw_blocks(total), 256, 0, as_stream(stream), reinterpret_cast<const float*>(frames), N, H, W, g, rh, rw, w_begin, w_count, reinterpret_cast<float*>(out));
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int mmr_window_blend(const float* win_logits, int N, int C, int H, int W, const int* ys, int ny,
                                const int* xs, int nx, int rh, int rw, float* out, int64_t* pred,
                                mmr_stream_t stream) {
  MMR_REQUIRE(win_logits && (out || pred) && ys && xs && C >= 1, "bad argument");
  WindowGrid g;
  if (fill_grid(g, ys, ny, xs, nx, H, W, rh, rw)) return -1;
  mmr_launch((window_blend_kernel), sw_blocks((int64_t)N * H * W), 256, 0, as_stream(stream), win_logits, N, C, H, W, g, rh, rw,
                                                                                    out, pred);
  MMR_CUDA_CHECK(cudaGetLastError());
  return 0;
}
