/*
 * mmrseg.h — C-ABI of libmmrseg.so: the sm_100a kernels behind the segmentation
 * train / infer step of AliakbarMzadeh/MMR_semantic-segmentation_v1.
 *
 * The reference has no FFI: its "operator interface" for this path is the set of
 * Python call sites listed below.  Every entry point names the
 * reference call it replaces.  Conventions:
 *   - plain pointers and sizes, no torch types; all device pointers are caller-owned;
 *   - every launch is asynchronous on the given CUDA stream (a cudaStream_t passed as
 *     void*); nothing here synchronises or allocates on the launch path (plan creation
 *     allocates its small descriptor tables once);
 *   - return 0 on success, a negative code on error; mmr_last_error() gives the
 *     message (thread-local);
 *   - activations are NHWC bf16 inside the path, logits / dlogits are NCHW fp32 at the
 *     boundary (what `model(x)` returns in the reference), labels are int64.
 */
#ifndef MMRSEG_H
#define MMRSEG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* mmr_stream_t; /* cudaStream_t */

const char* mmr_last_error(void);
int mmr_abi_version(void);
/* 1 if a CUDA device with compute capability 10.x is usable in this process. */
int mmr_device_ok(void);

/* ------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 (fprop and dgrad share it).
 * Replaces: nn.Conv2d forward/backward-data as issued by smp Conv2dReLU / torchvision
 * BasicBlock / SegmentationHead inside `seg = model(img)` (SU/ModelTraining.py:589,
 * SU/ModelEval.py:419, ED/Main_MMR_SegModel.py:697) and `loss.backward()`
 * (SU/ModelTraining.py:614, ED/Main_MMR_SegModel.py:715), including the
 * F.interpolate(nearest x2) + torch.cat of smp DecoderBlock.forward, which become
 * K-segments of the gather (no upsampled / concatenated tensor is materialised).
 *
 * One output tile = 128 pixels (box_w x box_h x box_n) x bn channels.  The reduction is a
 * list of K-steps; K-step i multiplies a [128 x bk] activation box, fetched by TMA from
 * source `src` at channel c0 and spatial origin (ax*gx0+bx, ay*gy0+by) of the tile, by the
 * [bn x bk] weight block at column wk of the K-major weight matrix.  Up to 4 pixel
 * classes (output parities) each own a K-step range, which is how nearest-x2 gathers and
 * stride-2 transposed gathers are expressed without replication or zero-insertion.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  const void* ptr; /* bf16 NHWC */
  int32_t C, W, H, N;
  int32_t es; /* TMA element stride along W and H (1 or 2) */
} MmrSrc;

typedef struct {
  int32_t src, c0, ax, bx, ay, by, wk, pad_;
} MmrKStep;

typedef struct {
  int32_t kbegin, kcount, oy_add, ox_add;
} MmrConvClass;

typedef struct {
  void* ptr;    /* destination tensor base (bf16 NHWC, or fp32 NCHW in out_mode 1) */
  int32_t ldc;  /* channels of the destination tensor */
  int32_t coff; /* first destination channel written by this N-tile */
  /* halo conv store groups only (0 / 1 = dense): the conv's pixel (y, x) is pixel (step*y + oy, step*x + ox) of
   * a destination tensor [N][step*H][step*W][ldc] -- depth-to-space through the TMA store's strides (the
   * space-to-depth stem, mmr_stem_s2d_*). */
  int32_t step, oy, ox;
  /* step == -2 (data-gradient launches, plain epilogue): the group is stored 2x2 SUM-POOLED into a tensor
   * [N][H/2][W/2][ldc] -- the gradient with respect to a nearest-x2 upsampled source (smp DecoderBlock). */
} MmrOutSeg;

enum { MMR_OUT_BF16_NHWC = 0, MMR_OUT_F32_NCHW = 1 };

typedef struct {
  int32_t nsrc;
  MmrSrc src[6];
  const void* weights; /* bf16 [w_rows][w_cols], K contiguous */
  int32_t w_rows, w_cols;
  int32_t bk; /* channels per K-step: 64, 32 or 16 */
  int32_t bn; /* output channels per tile: multiple of 16, <= 256 */
  int32_t box_w, box_h, box_n; /* box_w*box_h*box_n == 128 */
  int32_t ncls;
  MmrConvClass cls[4];
  int32_t nksteps;
  const MmrKStep* ksteps; /* host array */
  int32_t n_tiles_n;
  const MmrOutSeg* outsegs; /* host array, one per N-tile */
  int32_t gx_count, gy_count, n_img; /* extent of the tile grid (per class) */
  int32_t oy_mul, ox_mul;            /* output y = oy_mul*gy + oy_add */
  int32_t Hout, Wout;                /* spatial size of the destination tensors */
  int32_t cout_total;                /* valid output channels (masks the last N-tile) */
  const float* scale;                /* per output channel, may be NULL (=1) */
  const float* bias;                 /* per output channel, may be NULL (=0) */
  const void* residual;              /* bf16 NHWC, added before ReLU, may be NULL */
  int32_t res_ldc;
  int32_t relu;
  int32_t out_mode;
} MmrConvDesc;

int mmr_conv_plan_create(const MmrConvDesc* desc, void** plan);
/* impl: 0 = tcgen05/TMA kernel (the product path); 1 = scalar CUDA-core kernel reading the
 * same K-step tables (test aid for bisecting table bugs from tensor-core bugs). */
int mmr_conv_plan_run(void* plan, int impl, mmr_stream_t stream);
int mmr_conv_plan_destroy(void* plan);

/* ------------------------------------------------------------------------------------
 * 3x3 / stride 1 / pad 1 convolution, second kernel generation (csrc/conv_halo.cu): one halo
 * tile per channel chunk feeds all nine filter taps as shifted shared-memory views, one weight
 * slot is shared by the tx M-tiles (16 rows x 8 pixels each) of a macro tile, nearest-x2 sources
 * are replicated by zero-stride TMA dimensions.  Same call sites as mmr_conv_plan_* above: the
 * Conv2dReLU / BasicBlock / SegmentationHead convolutions of `seg = model(img)` and the
 * data-gradient half of `loss.backward()`; stride-2, 1x1 and 7x7 convolutions stay on
 * mmr_conv_plan_*.
 * ------------------------------------------------------------------------------------ */
/* BatchNorm2d training-mode finalisation fused into the kernel that produced the statistics: the
 * last CTA to finish (ticket counter) turns the accumulated sums into mean / invstd / scale / shift,
 * updates the running statistics (unbiased variance, momentum) and num_batches_tracked, and re-arms
 * (zeroes) the statistics slots and the ticket.  Same arithmetic as mmr_bn_finalize. */
typedef struct {
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* running_mean;
  float* running_var;
  int64_t* num_batches_tracked;
  float* mean;
  float* invstd;
  float* scale;
  float* shift;
  int64_t count;    /* values per channel: N*H*W */
  uint32_t* ticket; /* device counter, zero before the first launch */
} MmrBnFinalize;

/* BatchNorm backward reduction fused into the data-gradient launch that produces the ONLY gradient contribution dx
 * of a conv -> BatchNorm -> ReLU unit (no residual): the epilogue reads the unit's z tile, recomputes the ReLU
 * mask (z*mask_scale + mask_shift > 0), accumulates sum(g) and sum(g*z) of g = mask ? dx : 0 per channel, and the
 * last CTA (ticket) writes dgamma / dbeta / coef exactly as mmr_bn_bwd_reduce_fused does.  dx itself is stored
 * unmasked; mmr_bn_bwd_apply_masked recomputes the mask.  Needs one N tile = one store group = the whole tensor. */
typedef struct {
  const void* z;            /* bf16 NHWC, same shape as dx */
  const float* mask_scale;  /* the forward pass's scale / shift of the unit */
  const float* mask_shift;
  const float* mean;
  const float* invstd;
  const float* gamma;       /* may be NULL (= 1) */
  float* dgamma;            /* may be NULL */
  float* dbeta;             /* may be NULL */
  float* coef;              /* [3][C] */
  double* slots;            /* [8][2][C], zero before the first launch, re-armed by the last CTA */
  uint32_t* ticket;
  int64_t count;            /* N*H*W */
  int32_t accumulate;
} MmrBnBwdFused;

typedef struct {
  const void* ptr; /* bf16 NHWC, stored resolution (half of the conv's when up == 2) */
  int32_t C, W, H, N;
  int32_t up; /* 1, or 2 = nearest x2 (smp DecoderBlock F.interpolate) */
} MmrHaloSrc;

typedef struct {
  int32_t nsrc;
  MmrHaloSrc src[6];   /* channel-concatenated in this order (torch.cat of DecoderBlock.forward) */
  int32_t N, H, W;     /* resolution of the convolution (= output) */
  const void* weights; /* bf16, packed by mmr_pack_weights_halo */
  int32_t cb;          /* channels per K chunk: 64, 32 or 16; divides every source's C */
  int32_t bn;          /* output channels per N tile: multiple of 16, <= 256 */
  int32_t sg;          /* channels per store group (one TMA store): 16, 32 or 64, divides bn */
  int32_t n_ntiles;
  int32_t tx;          /* M-tiles per macro tile: 1, 2 or 4 */
  int32_t tps;         /* filter taps per weight slot: 1, 3 or 9; tps*bn <= 256 */
  int32_t halo_stages, w_slots, acc_bufs, out_stages; /* shared-memory / TMEM pipeline depths */
  int32_t direct_store; /* 1: store bf16 tiles from registers (narrow groups) instead of TMA */
  int32_t ngroups;     /* n_ntiles * bn/sg store groups (bf16 NHWC mode) */
  const MmrOutSeg* groups; /* host array: destination tensor, its channel count, first channel */
  int32_t cout_total;  /* valid output channels */
  const float* scale;  /* per output channel, may be NULL */
  const float* bias;
  const void* residual; /* bf16 NHWC [N][H][W][res_ldc], indexed by output channel, may be NULL */
  int32_t res_ldc, relu;
  int32_t out_mode;    /* MMR_OUT_BF16_NHWC, or MMR_OUT_F32_NCHW into out_f32 (one N tile) */
  void* out_f32;
  int32_t out_ldc;
  /* optional [8][2][stats_ld] doubles, ACCUMULATED: per-channel sum and sum of squares of the
   * stored bf16 outputs (BatchNorm batch statistics); consumed by mmr_bn_finalize(nblk = 8). */
  double* stats;
  int32_t stats_ld;
  const MmrBnFinalize* bn_finalize; /* optional (needs stats): fused mmr_bn_finalize */
  /* Row-phase stacking: 1 (off), 2 or 4.  An M tile takes every rph-th output row and one MMA of
   * N = (phases) * bn serves up to three vertically adjacent output rows from ONE read of the activation
   * operand (what bounds N <= 64 MMAs).  Needs tps = 3, bn <= 64, H % rph == 0 (H % (16 rph) == 0 with
   * nearest-x2 sources), weights packed with layout 1, tx * rph * bn * acc_bufs <= 512. */
  int32_t rph;
  const MmrBnBwdFused* bn_bwd; /* optional (data-gradient launches): see MmrBnBwdFused */
  const struct MmrHeadMetric* head_metric; /* optional (fp32 NCHW head launches): see MmrHeadMetric */
  /* halo tile loader: 0 = one TMA box per chunk; 1 = cp.async gather by two warps (cb <= 32 only, halo_stages >= 2):
   * TMA moves 32- / 64-byte pixel rows one at a time (~4 clk per row) and bounds the 16- / 32-channel 512^2 layers */
  int32_t loader;
} MmrHaloConvDesc;

/* Eval-time metric fused into the segmentation head's epilogue: the reference's
 *   seg = model(img); evaluator.addBatch(seg, oneHotGT, args); seg = torch.argmax(seg, 1)
 * (SU/ModelTraining.py:736-760, SU/ModelEval.py:363-458, SU/utils.py:109-133) without the logits ever reaching
 * HBM.  Per pixel, on the fp32 logits the epilogue holds in registers (accumulator + bias, the very values the
 * plain head launch would store): pred = first maximal class (torch.argmax's rule, NaN counts as maximal);
 * pred_out[n][y][x] = pred (uint8, optional); confusion[n][label][pred] += 1 (unsigned 64-bit, counted with
 * shared-memory 32-bit integer atomics per CTA and image, optional; labels outside [0, classes) are skipped as
 * in mmr_confusion_from_logits).  With a head_metric the descriptor's out_f32 may be NULL (no logits written). */
typedef struct MmrHeadMetric {
  const void* labels;          /* [N][H][W], int64 (labels_u8 = 0) or uint8 (labels_u8 = 1); NULL: no confusion */
  int32_t labels_u8;
  uint8_t* pred_out;           /* [N][H][W] or NULL */
  unsigned long long* confusion; /* [N][classes][classes], accumulated, or NULL */
} MmrHeadMetric;

int mmr_halo_conv_plan_create(const MmrHaloConvDesc* desc, void** plan);
int mmr_halo_conv_plan_run(void* plan, mmr_stream_t stream);
int mmr_halo_conv_plan_destroy(void* plan);
/* Diagnostic (MMR_HALO_DBG bit 16): per-CTA %globaltimer stamps of the LAST mmr_halo_conv_plan_run launch, eight
 * per CTA (entry, after the grid dependency wait, after setup, first operands landed, last MMA committed, last
 * item stored, finalisation done, exit), copied to host memory.  Synchronises the device; never on the hot path. */
int mmr_debug_halo_trace(unsigned long long* out_host, int n_ctas);
/* OIHW fp32 3x3 master weights -> bf16 [n_ntiles][nchunks][9][bn][cb].  mode 0 (fprop): rows are
 * output channels, columns input channels; mode 1 (dgrad): rows are input channels, columns output
 * channels, taps mirrored.  Out-of-range rows / columns are zero.  layout: see MmrPackJob. */
int mmr_pack_weights_halo(const float* w_oihw, int O, int I, int mode, int cb, int bn, int n_ntiles,
                          int nchunks, int layout, void* out, mmr_stream_t stream);
/* The same for every layer of a network in one launch: jobs_dev is a DEVICE array of njobs jobs
 * (the optimiser rewrites the fp32 masters every step, so the packing runs once per step). */
typedef struct {
  const float* w_oihw;
  void* out;
  int32_t O, I, mode, cb, bn, n_ntiles, nchunks;
  int32_t layout; /* order of the 9 taps inside a chunk: 0 = ky*3+kx; 1 = [kx][ky = 2,1,0] (rph > 1) */
} MmrPackJob;
/* total_blocks = sum over the jobs of ceil(n_ntiles * nchunks * bn * cb / mmr_pack_items_per_block()) (the host
 * built the table). */
int mmr_pack_items_per_block(void);
int mmr_pack_weights_halo_batch(const MmrPackJob* jobs_dev, int njobs, int64_t total_blocks, mmr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Weight gradient on tcgen05.  Replaces the weight-gradient half of `loss.backward()`
 * (SU/ModelTraining.py:614, ED/Main_MMR_SegModel.py:715) for every nn.Conv2d on the path.
 *
 *   dWt[(chunk, ch)][co] = sum over pixels  x[chunk.src][pixel + shift(chunk)][chunk.c0 + ch] * dz[pixel][co]
 *
 * Both operands are pixel-major in memory (NHWC), i.e. MN-major UMMA operands: a K-step is a
 * box of kp_w*kp_h*kp_n = 32 pixels; the M side stacks 128/chunk_ch activation chunks (a
 * chunk = source, channel offset, filter tap), the N side is the conv's output channels.  A
 * CTA keeps up to 512/cout accumulator tiles in TMEM so one dz box is reused by all of them.
 * The pixel range is split n_split ways (split-K); fp32 partials are then summed in a fixed
 * order (deterministic) and scattered into the OIHW fp32 gradient.
 * Up to 4 pixel classes (output parities) carry per-class box shifts, as in the forward plan.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  int32_t src, c0;      /* source tensor, first channel of the chunk */
  int32_t a;            /* box origin = a*g0 + b[cls] along x and y */
  int32_t bx[4], by[4]; /* per pixel class */
  int32_t dst_ci;       /* input-channel index of channel 0 of the chunk in the OIHW gradient */
  int32_t dst_tap;      /* filter tap (ky*kw + kx); -1 marks a padding chunk (discarded) */
} MmrWgChunk;

typedef struct {
  MmrSrc dz; /* gradient wrt the conv output, bf16 NHWC (C = padded channel count) */
  int32_t dz_a;
  int32_t dz_bx[4], dz_by[4];
  int32_t nsrc;
  MmrSrc src[6];
  int32_t ncls;
  int32_t nchunks;          /* multiple of 128/chunk_ch */
  const MmrWgChunk* chunks; /* host array */
  int32_t chunk_ch;         /* channels per chunk: 64, 32 or 16 */
  int32_t cout_gemm;        /* N of the GEMM: dz channels used, multiple of 16 */
  int32_t kp_w, kp_h, kp_n; /* pixel box of one K-step, product 32 */
  int32_t gx_count, gy_count, n_img;
  int32_t n_split;
  float* partial; /* fp32 [n_split][nchunks*chunk_ch][cout_gemm] */
  float* dst;     /* OIHW fp32 gradient [dst_cout][dst_cin][dst_taps] */
  int32_t dst_cout, dst_cin, dst_taps;
  int32_t chunk_valid_ch; /* leading channels of each chunk that exist in dst (<= chunk_ch) */
} MmrWgradDesc;

int mmr_wgrad_plan_create(const MmrWgradDesc* desc, void** plan);
/* impl: 0 = tcgen05 kernel, 1 = scalar CUDA-core kernel on the same tables (test aid).
 * accumulate != 0 adds into dst (gradient accumulation, ED/Main_MMR_SegModel.py:718). */
int mmr_wgrad_plan_run(void* plan, int impl, int accumulate, mmr_stream_t stream);
int mmr_wgrad_plan_destroy(void* plan);

/* ------------------------------------------------------------------------------------
 * Weight gradient of a 3x3 / stride 1 / pad 1 convolution, second kernel generation
 * (csrc/conv_wgrad_halo.cu): the nine shifted activation boxes of a pixel tile are views of one
 * halo tile; a CTA owns (channel chunk, bn-wide output-channel slice) x a split of the pixel macro
 * tiles; partials [nchunks*cout_gemm/bn][n_split][A*128][bn] fp32 (A = 5 for cb 64, else 3) are
 * summed in split order and scattered to the OIHW fp32 gradient.  Same call site as
 * mmr_wgrad_plan_*: the weight-gradient half of `loss.backward()`.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  MmrHaloSrc dz; /* bf16 NHWC gradient wrt the conv output, C >= cout_gemm, up = 1 */
  int32_t nsrc;
  MmrHaloSrc src[6];
  int32_t N, H, W;
  int32_t cb;        /* channels per chunk: 64/32/16, divides every source's C */
  int32_t bn;        /* output channels per CTA slice: 64/32/16 */
  int32_t cout_gemm; /* dz channels used, multiple of bn */
  int32_t tx;        /* pixel tiles (16 rows x 8 pixels) per macro tile: 1, 2 or 4 */
  int32_t n_split;
  float* partial;    /* mmr_wgrad_halo_partial_floats(...) floats */
  float* dst;        /* OIHW fp32 [dst_cout][dst_cin][3][3] */
  int32_t dst_cout, dst_cin;
  int32_t mode;      /* 0: the nine taps are views of the activation halo tile (any cb);
                        1: third generation, cb = 64, bn = 64 / 32, tx <= 2: the filter column is a one-pixel shift of
                           the dz tile (N = 3 bn per tcgen05.mma); partials [slices][n_split][192][3 bn];
                        2: the same formulation for narrow layers, cb = 16 / 32, bn = 16 / 32, tx = 2 / 4, tiles gathered
                           with cp.async by six warps; partials [slices][n_split][3 cb][3 bn] */
} MmrWgradHaloDesc;

int64_t mmr_wgrad_halo_partial_floats(int nchunks, int cb, int bn, int n_ntiles, int n_split);
int64_t mmr_wgrad_kx_partial_floats(int nchunks, int bn, int n_ntiles, int n_split); /* mode 1 */
int64_t mmr_wgrad_thin_partial_floats(int nchunks, int cb, int bn, int n_ntiles, int n_split); /* mode 2 */
int mmr_wgrad_halo_plan_create(const MmrWgradHaloDesc* desc, void** plan);
/* accumulate != 0 adds into dst (gradient accumulation). */
int mmr_wgrad_halo_plan_run(void* plan, int accumulate, mmr_stream_t stream);
int mmr_wgrad_halo_plan_destroy(void* plan);

/* ------------------------------------------------------------------------------------
 * Layout / packing kernels at the boundary of the path.
 * ------------------------------------------------------------------------------------ */
/* NCHW fp32 image -> im2col rows for the 7x7 stride-2 pad-3 stem (encoder.conv1), bf16
 * [N*Ho*Wo][kpad], column = c*49 + ky*7 + kx (the OIHW order of conv1.weight[o]), zero padded to kpad
 * (152 <= kpad <= 256).  Optionally applies (x-mean[c])/std[c] first (utils.normalize,
 * SU/utils.py:480-519) when mean != NULL. */
int mmr_stem_im2col(const float* x, int N, int H, int W, void* out, int kpad, const float* mean,
                    const float* std_, mmr_stream_t stream);
/* The same from uint8 HWC frames [N][H][W][3] (what the data loaders hold before ToTensor,
 * SU/SegNetDataLoaderV1_SAR.py:150-151): x/255, then the optional normalisation.  Replaces the CPU
 * ToTensor + utils.normalize + fp32 host-to-device copy in front of the model. */
int mmr_stem_im2col_u8(const uint8_t* x_nhwc, int N, int H, int W, void* out, int kpad, const float* mean,
                       const float* std_, mmr_stream_t stream);
/* Space-to-depth form of the 7x7 stride-2 pad-3 stem (encoder.conv1) for the halo kernel: with 4x4 pixel blocks
 * as channels the stem is a 3x3 pad-1 convolution from 48 (+16 zero) block channels to 4 output phases x Cout
 * channels on the H/4 x W/4 grid (phase (qy, qx) of block (Y, X) is output pixel (2Y + qy, 2X + qx)):
 *   input row 2*oy + ky - 3 = 4*(Y + A - 1) + ry   <=>   ky = 4*(A - 1) + ry - 2*qy + 3,  A in {0,1,2}.
 * No im2col matrix (320 B per output pixel, written and read back) exists; the blocked image is 6 B per input
 * pixel.  mmr_stem_s2d_pack: image -> bf16 [N][H/4][W/4][64], channel (ry*4 + rx)*3 + c; fp32 NCHW input, or
 * (is_u8) uint8 NHWC frames scaled by 1/255; optional (x - mean[c]) / std[c].
 * mmr_stem_s2d_weights: conv1.weight fp32 [Cout][3][7][7] -> fp32 OIHW [4*Cout][64][3][3] (row = phase*Cout + co),
 * which mmr_pack_weights_halo then packs like any 3x3 layer. */
int mmr_stem_s2d_pack(const void* x, int is_u8, int N, int H, int W, void* out, const float* mean,
                      const float* std_, mmr_stream_t stream);
int mmr_stem_s2d_weights(const float* w7, int Cout, float* w3_oihw, mmr_stream_t stream);
/* Backward of mmr_stem_s2d_weights: dw7[co][c][ky][kx] (+)= sum over the four phases of the matching element of
 * dw3 [4*Cout][64][3][3] (the weight gradient of the 3x3 form, from mmr_wgrad_halo_plan_* with a phased dz:
 * MmrWgradHaloDesc.dz.up = 2 means dz is [N][2H][2W][C] and GEMM output channel (q, c) is channel c of pixel
 * (2y + qy, 2x + qx); needs bn = C, cout_gemm = 4 C). */
int mmr_stem_s2d_wgrad_fold(const float* dw3_oihw, int Cout, float* dw7, int accumulate, mmr_stream_t stream);
/* NCHW fp32 -> NHWC bf16 with channel padding to cpad (zeros). */
int mmr_pack_nchw_f32_to_nhwc_bf16(const float* x, int N, int C, int H, int W, void* out,
                                   int cpad, mmr_stream_t stream);
/* uint8 HWC frames [N][H][W][3] -> NHWC bf16 padded to cpad channels: x/255, then (x-mean)/std when
 * mean != NULL (the image feed of the 3x3 full-resolution convs of ResNetUNet). */
int mmr_pack_nhwc_u8_to_nhwc_bf16(const uint8_t* x, int N, int H, int W, void* out, int cpad,
                                  const float* mean, const float* std_, mmr_stream_t stream);
/* NHWC bf16 -> NCHW fp32 (first C of ldc channels). */
int mmr_unpack_nhwc_bf16_to_nchw_f32(const void* x, int N, int C, int ldc, int H, int W,
                                     float* out, mmr_stream_t stream);
/* OIHW fp32 master weights -> bf16 GEMM layouts:
 *   fwd  [O][kh*kw][I]  (row stride ldf >= kh*kw*I)   and, if dgrad != NULL,
 *   dgrad [I][kh*kw][o_pad] (row stride ldd >= kh*kw*o_pad; o_pad >= O, pad columns untouched). */
int mmr_repack_weights(const float* w_oihw, int O, int I, int taps, void* fwd, int ldf,
                       void* dgrad, int ldd, int o_pad, mmr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * BatchNorm2d (+ residual add + ReLU), training and eval; replaces nn.BatchNorm2d / ReLU
 * / the BasicBlock residual add inside model(img) and their backward.
 * ------------------------------------------------------------------------------------ */
/* Per-channel sum and sum of squares of z [P][C] bf16 -> double partial[nblk][2][C]. */
int mmr_bn_stats(const void* z, int64_t P, int C, double* partial, int nblk, mmr_stream_t stream);
/* partial -> mean/invstd, scale = gamma*invstd, shift = beta - mean*scale; running stats
 * updated with `momentum` (unbiased variance), num_batches_tracked += 1. */
int mmr_bn_finalize(const double* partial, int nblk, int64_t P, int C, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean,
                    float* running_var, int64_t* num_batches_tracked, float* mean, float* invstd,
                    float* scale, float* shift, mmr_stream_t stream);
/* a = [relu](z*scale + shift [+ residual]) ; all NHWC bf16, [P][C]. */
int mmr_bn_apply(const void* z, int64_t P, int C, const float* scale, const float* shift,
                 const void* residual, int relu, void* out, mmr_stream_t stream);

/* A gradient of an activation is the sum of several contribution tensors (one per
 * consumer).  pool2 != 0: the contribution lives at twice the resolution and is 2x2
 * sum-pooled on the fly (backward of nearest x2 upsampling). */
typedef struct {
  const void* ptr; /* bf16 NHWC */
  int32_t pool2;
} MmrContrib;

/* g = mask * sum(contribs), mask = (act > 0) if act != NULL; writes g (bf16) and the
 * per-channel double partial sums of g and g*xhat, xhat = (z-mean)*invstd. */
int mmr_bn_bwd_reduce(const MmrContrib* contribs, int ncontrib, const void* act, const void* z,
                      const float* mean, const float* invstd, int N, int H, int W, int C, void* g,
                      double* partial, int nblk, mmr_stream_t stream);
/* mmr_bn_bwd_reduce + mmr_bn_bwd_finalize in one launch: block sums go to slots [8][2][C] (double
 * atomics, zero before the first launch), the last CTA (ticket) writes dgamma / dbeta / coef and
 * re-arms slots and ticket.  ReLU mask: (act > 0) when act != NULL, or -- for units without a
 * residual -- (z * mask_scale + mask_shift > 0) with the forward pass's scale / shift, which is the
 * same predicate without reading the activation back; neither: no mask. */
int mmr_bn_bwd_reduce_fused(const MmrContrib* contribs, int ncontrib, const void* act, const void* z,
                            const float* mean, const float* invstd, int N, int H, int W, int C, void* g,
                            double* slots, int nblk, const float* gamma, float* dgamma, float* dbeta,
                            int accumulate, float* coef, uint32_t* ticket, const float* mask_scale,
                            const float* mask_shift, mmr_stream_t stream);
/* partial -> dgamma, dbeta (fp32, accumulate flag) and the coefficients used by apply. */
int mmr_bn_bwd_finalize(const double* partial, int nblk, int64_t P, int C, const float* gamma,
                        const float* invstd, float* dgamma, float* dbeta, int accumulate,
                        float* coef /* [3][C] */, mmr_stream_t stream);
/* dz = coefA*g + coefB*xhat + coefC, bf16. */
int mmr_bn_bwd_apply(const void* g, const void* z, const float* mean, const float* invstd,
                     const float* coef, int64_t P, int C, void* dz, mmr_stream_t stream);
/* The same for a unit with ONE full-resolution gradient contribution dx and no residual: the masked
 * gradient g = (z*mask_scale + mask_shift > 0) ? dx : 0 is recomputed here, so mmr_bn_bwd_reduce_fused
 * may be called with g = NULL (it then writes nothing but the sums). */
int mmr_bn_bwd_apply_masked(const void* dx, const void* z, const float* mean, const float* invstd,
                            const float* coef, const float* mask_scale, const float* mask_shift, int64_t P,
                            int C, void* dz, mmr_stream_t stream);
/* g = mask * sum(contribs) only (layers without BN), plus optional per-channel sum (bias
 * gradient) as double partials. */
int mmr_grad_gather(const MmrContrib* contribs, int ncontrib, const void* act, int N, int H, int W,
                    int C, void* g, double* partial, int nblk, mmr_stream_t stream);

/* dbias[c] (+)= sum over blocks of the per-channel sums mmr_grad_gather left in partial. */
int mmr_bias_grad_finalize(const double* partial, int nblk, int C, float* dbias, int accumulate,
                           mmr_stream_t stream);

/* MaxPool2d(3, stride 2, pad 1) on NHWC bf16; idx = window position 0..8 of the first
 * maximum (torch tie rule), may be NULL (eval mode: only the backward pass reads it).  Replaces encoder.maxpool. */
int mmr_maxpool3x3s2_fwd(const void* x, int N, int H, int W, int C, void* out, uint8_t* idx,
                         mmr_stream_t stream);
int mmr_maxpool3x3s2_bwd(const MmrContrib* contribs, int ncontrib, const uint8_t* idx, int N,
                         int H, int W, int C, void* gin, mmr_stream_t stream);
/* nn.MaxPool2d(2) of the in-tree UNet's Down block (SU/UArchModel/unet_parts.py): disjoint 2x2 windows,
 * floor mode; idx (uint8, 0..3) = position of the first maximum in scan order (may be NULL). */
int mmr_maxpool2x2s2_fwd(const void* x, int N, int H, int W, int C, void* out, uint8_t* idx,
                         mmr_stream_t stream);
int mmr_maxpool2x2s2_bwd(const MmrContrib* contribs, int ncontrib, const uint8_t* idx, int N, int H, int W,
                         int C, void* gin, mmr_stream_t stream);

/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) on NHWC bf16 and its adjoint;
 * replaces self.upsample of the reference's ResNetUNet (SU/UArchModel/resnet_unet.py:195,262-294).
 * x: [N][H][W][C] -> out: [N][2H][2W][C].  The backward gathers the (summed) gradient
 * contributions of the upsampled tensor (each optionally 2x2-pooled) into gin [N][H][W][C]. */
int mmr_upsample_bilinear2x_fwd(const void* x, int N, int H, int W, int C, void* out,
                                mmr_stream_t stream);
int mmr_upsample_bilinear2x_bwd(const MmrContrib* contribs, int ncontrib, int N, int H, int W, int C,
                                void* gin, mmr_stream_t stream);

/* Deep-supervision heads (BASELINE config 4; no reference semantics, definition in
 * oracle/unetpp.py::DeepSupervisionUnetPlusPlus): fp32 NCHW logits [planes][h][w] of an auxiliary
 * head -> nearest-upsampled [planes][h*f][w*f], and the adjoint f x f sum-pool for the gradient. */
int mmr_upsample_nearest_f32_nchw(const float* in, int64_t planes, int h, int w, int f, float* out,
                                  mmr_stream_t stream);
int mmr_sumpool_f32_nchw(const float* in, int64_t planes, int h, int w, int f, float* out,
                         mmr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Loss: softmax + soft-Dice + cross-entropy, forward and backward.
 * Replaces dice_loss / DiceLoss.forward (SU/dice_loss.py:37-161,241-259) +
 * nn.CrossEntropyLoss (SU/ModelTraining.py:360,601) combined as w*dice + (1-w)*ce
 * (SU/ModelTraining.py:600-603), and monai DiceCELoss(softmax=True)
 * (ED/Main_MMR_SegModel.py:578,709) through its smoothing constants.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  float dice_eps_nr, dice_eps_dr; /* 1.0/1.0 (SU) or 1e-5/1e-5 (monai) */
  float onehot_eps;               /* 1e-6: kornia one_hot adds eps everywhere; 0 for monai */
  float w_dice, w_ce;
  int32_t dice_channels;   /* leading channels entering the dice term (ignore_index slice) */
  int64_t ce_ignore_index; /* labels equal to this are skipped by CE (-100 = torch default) */
} MmrLossParams;

/* workspace: doubles, size mmr_dice_ce_workspace_doubles(N, C, blocks). Outputs:
 * out[0] = total loss, out[1] = dice term, out[2] = ce term (fp32). */
int64_t mmr_dice_ce_workspace_doubles(int N, int C, int nblk);
int mmr_dice_ce_fwd(const float* logits, const int64_t* labels, int N, int C, int H, int W,
                    const MmrLossParams* p, double* workspace, int nblk, float* out,
                    mmr_stream_t stream);
/* dlogits (fp32 NCHW) = grad_scale * [*grad_scale_dev] * dLoss/dlogits, using the sums left in
 * workspace.  grad_scale_dev (device scalar, may be NULL) carries autograd's upstream gradient
 * without a host synchronisation. */
int mmr_dice_ce_bwd(const float* logits, const int64_t* labels, int N, int C, int H, int W,
                    const MmrLossParams* p, const double* workspace, float grad_scale,
                    const float* grad_scale_dev, float* dlogits, mmr_stream_t stream);
/* fp32 NCHW dlogits -> bf16 NHWC padded to cpad channels + per-class sums (head bias grad). */
int mmr_head_grad_prep(const float* dlogits, int N, int C, int H, int W, void* out, int cpad,
                       float* dbias, int accumulate, mmr_stream_t stream);

/* 1x1 classification head over a 64-channel bf16 NHWC feature map on the CUDA cores, in fp32 from the fp32
 * master weights w[Cout][64] (HBM-bound; nothing padded to a tensor-core tile, no repacked weights): replaces the
 * forward / data-gradient / weight-gradient of `ResNetUNet.conv_last = nn.Conv2d(64, n_class, 1)`
 * (SU/UArchModel/resnet_unet.py:204, :298).  Cin must be 64, 1 <= Cout <= 16.
 * fwd: logits fp32 NCHW [N][Cout][H][W] = bias + x . w^T.
 * bwd: ONE pass over x and dlogits (fp32 NCHW) writes dx bf16 NHWC (times the ReLU mask x > 0 when relu_mask: dx is
 * then the dz of the conv + bias + ReLU that produced x), dw[Cout][64] and dbias[Cout] of the head and
 * dbias_producer[64] = column sums of the stored dx (may be NULL); (+)= when accumulate.  workspace: floats for the
 * per-CTA partial sums, mmr_pointwise_head_bwd_workspace_bytes(Cout) bytes. */
int mmr_pointwise_head_fwd(const void* x, const float* w, const float* bias, int N, int H, int W, int Cin,
                           int Cout, float* logits, mmr_stream_t stream);
int64_t mmr_pointwise_head_bwd_workspace_bytes(int Cout);
int mmr_pointwise_head_bwd(const float* dlogits, const void* x, const float* w, int N, int H, int W, int Cin,
                           int Cout, int relu_mask, void* dx, float* dw, float* dbias, float* dbias_producer,
                           int accumulate, float* workspace, mmr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Sliding-window inference (monai.inferers.sliding_window_inference with mode="constant", called at
 * ED/Main_MMR_SegModel.py:1308-1317, followed by preds.argmax(1) at :1320).  Windows of one frame are
 * ordered (iy, ix) over the start lists ys[ny] x xs[nx]; window w = frame * ny*nx + iy*nx + ix.
 * ------------------------------------------------------------------------------------ */
/* Copy windows [w_begin, w_begin + w_count) into the predictor's input batch: fp32 NCHW frames
 * [N][3][H][W] -> [w_count][3][rh][rw], or (is_u8) uint8 NHWC frames [N][H][W][3] -> [w_count][rh][rw][3].
 * Windows past the last one (padding of the final batch) are zero. */
int mmr_window_gather(const void* frames, int is_u8, int N, int H, int W, const int* ys, int ny, const int* xs,
                      int nx, int rh, int rw, int w_begin, int w_count, void* out, mmr_stream_t stream);
/* out[n][c][y][x] = mean over the windows covering (y, x) of win_logits[w][c][y - y0(w)][x - x0(w)], summed in
 * window order by one thread per pixel (deterministic); pred (int64 [N][H][W], optional) = argmax over c of the
 * blended logits with torch's first-maximum / NaN rule.  Either output may be NULL. */
int mmr_window_blend(const float* win_logits, int N, int C, int H, int W, const int* ys, int ny, const int* xs,
                     int nx, int rh, int rw, float* out, int64_t* pred, mmr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Metric: argmax -> per-image confusion matrix, integer and bit-exact.
 * Replaces torch.argmax + Evaluate.addBatch (SU/utils.py:82-138) and smp get_stats
 * (ED/Main_MMR_SegModel.py:634-639,1323-1325): cm[n][g][p] += #{label==g & pred==p}.
 * Labels outside [0,C) (ignore_index) are skipped.  pred_out (int64 [N][H][W]) optional.
 * ------------------------------------------------------------------------------------ */
int mmr_confusion_from_logits(const float* logits, const int64_t* labels, int N, int C, int H,
                              int W, int64_t* cm /* [N][C][C], accumulated */, int64_t* pred_out,
                              mmr_stream_t stream);
/* ignore_index: labels equal to it are skipped (smp get_stats ignore_index).  overflow_bin != 0:
 * cm is [N][C+1][C+1] and bin C collects out-of-range predictions / labels (smp counts a pixel
 * whose prediction is out of range as a false negative of its label); otherwise such pixels
 * are skipped and cm is [N][C][C]. */
int mmr_confusion_from_preds(const int64_t* preds, const int64_t* labels, int N, int C,
                             int64_t npix_per_image, int64_t ignore_index, int overflow_bin,
                             int64_t* cm, mmr_stream_t stream);
/* One-hot map [N][C][H][W] (fp32 when is_float, else int64) -> labels [N][H][W] int64 (first
 * maximal channel).  The reference passes one-hot ground truth to Evaluate.addBatch
 * (SU/ModelTraining.py:725,757) and to DiceCELoss (ED/Main_MMR_SegModel.py:700-709). */
int mmr_onehot_to_labels(const void* onehot, int is_float, int N, int C, int H, int W,
                         int64_t* labels, mmr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Per-image, per-class Hausdorff distance of the predicted vs the labelled mask.  Replaces the
 * reference's every-25-epochs loop: one_hot -> .cpu() -> skimage.metrics.hausdorff_distance(seg_slice,
 * label_slice) per image and class (SU/ModelTraining.py:625-649, 765-789).
 *   hd2[n][c] = max( max_{b: label == c} min_{a: pred == c} |a - b|^2 , the same with the roles swapped )
 * as an exact integer (separable Euclidean distance transform); 0 when both masks are empty, ~0ull (all ones)
 * when exactly one is (skimage: inf).  The host takes sqrt in float64.  pred: [N][H][W] uint8 (pred_u8 = 1) or
 * int64; labels int64; workspace: >= mmr_hausdorff_workspace_bytes(C, H, W) (one image; more lets several images
 * share a launch).  classes <= 16.
 * ------------------------------------------------------------------------------------ */
int64_t mmr_hausdorff_workspace_bytes(int C, int H, int W);
int mmr_hausdorff_sq(const void* pred, int pred_u8, const int64_t* labels, int N, int C, int H, int W,
                     void* workspace, int64_t workspace_bytes, unsigned long long* hd2, mmr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Optimiser: Adam / AdamW over one flat fp32 buffer.  Replaces optim.Adam(...).step()
 * (SU/ModelTraining.py:366,617) and AdamW (ED/Main_MMR_SegModel.py:878-880).
 * mode 0: L2 (g += wd*p), mode 1: decoupled (AdamW).  bc1 = 1-b1^t, bc2 = 1-b2^t.
 * grad_scale multiplies g first (1/world_size for DDP, or the clip coefficient).
 * ------------------------------------------------------------------------------------ */
int mmr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1,
                  float b2, float eps, float wd, float bc1, float bc2, int mode, float grad_scale,
                  mmr_stream_t stream);
/* cudaMemsetAsync(ptr, 0, nbytes) on the stream: re-arms accumulated buffers (statistics slots of
 * mmr_halo_conv_plan_*, split-K partials) inside a replayed launch list. */
int mmr_zero_async(void* ptr, int64_t nbytes, mmr_stream_t stream);
/* torch.optim.SGD(lr, momentum, weight_decay) with dampening 0, no Nesterov (SU/ModelTraining.py:372,381):
 * g' = g*grad_scale + wd*p; buf = first_step ? g' : momentum*buf + g'; p -= lr*buf.  momentum_buf may be NULL
 * when momentum == 0. */
int mmr_sgd_step(float* p, const float* g, float* momentum_buf, int64_t n, float lr, float momentum, float wd,
                 int first_step, float grad_scale, mmr_stream_t stream);
/* sum of squares of g (double out[0] accumulated) for clip_grad_norm_. */
int mmr_sumsq(const float* g, int64_t n, double* out, mmr_stream_t stream);
/* Second half of torch.nn.utils.clip_grad_norm_(parameters, max_norm) (ED/Main_MMR_SegModel.py:722):
 * total = sqrt(sumsq[0]); coef = min(1, max_norm / (total + 1e-6)); g *= coef when coef < 1, in place.  The
 * sum of squares stays on the device (no host read); norm_out (may be NULL) receives total as fp32. */
int mmr_clip_scale(float* g, int64_t n, const double* sumsq, float max_norm, float* norm_out, mmr_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Eval-mode BatchNorm folding for every conv of a network in ONE launch (model.eval() forward,
 * SU/ModelTraining.py:703,736; SU/ModelEval.py:363-458): for each job
 *   scale[r*C + c] = gamma[c] * rsqrt(running_var[c] + eps)
 *   shift[r*C + c] = beta[c] - (running_mean[c] - conv_bias[c]) * scale[c]        r = 0 .. rep-1
 * conv_bias may be NULL (bias-free conv in front of the BatchNorm); rep > 1 repeats the layer's channels
 * (the space-to-depth stem's four output phases).  jobs_dev is a DEVICE array.  Runs inside the replayed
 * eval launch list, so the folded constants always follow the current parameters and running statistics.
 * ------------------------------------------------------------------------------------ */
typedef struct {
  const float* gamma;
  const float* beta;
  const float* running_mean;
  const float* running_var;
  const float* conv_bias;
  float* scale;
  float* shift;
  int32_t C, rep;
} MmrBnFoldJob;
int mmr_bn_fold_batch(const MmrBnFoldJob* jobs_dev, int njobs, float eps, mmr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMRSEG_H */
