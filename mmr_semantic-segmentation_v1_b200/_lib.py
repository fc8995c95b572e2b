"""ctypes binding of libmmrseg.so (the C-ABI declared in include/mmrseg.h).

There is no CPU fallback: if the shared library is missing this module raises, and every
launch wrapper raises `MmrError` carrying `mmr_last_error()` when a call fails.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MMR_LIB: A/B measurements against a variant build (mmrseg_b200.build.build_variant); same C-ABI, still no fallback
LIB_PATH = os.environ.get("MMR_LIB") or os.path.join(HERE, "libmmrseg.so")


class MmrError(RuntimeError):
    pass


class MmrSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
                ("N", C.c_int32), ("es", C.c_int32)]


class MmrKStep(C.Structure):
    _fields_ = [("src", C.c_int32), ("c0", C.c_int32), ("ax", C.c_int32), ("bx", C.c_int32),
                ("ay", C.c_int32), ("by", C.c_int32), ("wk", C.c_int32), ("pad_", C.c_int32)]


class MmrConvClass(C.Structure):
    _fields_ = [("kbegin", C.c_int32), ("kcount", C.c_int32), ("oy_add", C.c_int32),
                ("ox_add", C.c_int32)]


class MmrOutSeg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ldc", C.c_int32), ("coff", C.c_int32),
                ("step", C.c_int32), ("oy", C.c_int32), ("ox", C.c_int32)]


class MmrConvDesc(C.Structure):
    _fields_ = [
        ("nsrc", C.c_int32), ("src", MmrSrc * 6),
        ("weights", C.c_void_p), ("w_rows", C.c_int32), ("w_cols", C.c_int32),
        ("bk", C.c_int32), ("bn", C.c_int32),
        ("box_w", C.c_int32), ("box_h", C.c_int32), ("box_n", C.c_int32),
        ("ncls", C.c_int32), ("cls", MmrConvClass * 4),
        ("nksteps", C.c_int32), ("ksteps", C.POINTER(MmrKStep)),
        ("n_tiles_n", C.c_int32), ("outsegs", C.POINTER(MmrOutSeg)),
        ("gx_count", C.c_int32), ("gy_count", C.c_int32), ("n_img", C.c_int32),
        ("oy_mul", C.c_int32), ("ox_mul", C.c_int32),
        ("Hout", C.c_int32), ("Wout", C.c_int32),
        ("cout_total", C.c_int32),
        ("scale", C.c_void_p), ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("res_ldc", C.c_int32),
        ("relu", C.c_int32), ("out_mode", C.c_int32),
    ]


class MmrBnFinalize(C.Structure):
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("momentum", C.c_float),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
                ("mean", C.c_void_p), ("invstd", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("count", C.c_int64), ("ticket", C.c_void_p)]


class MmrBnBwdFused(C.Structure):
    _fields_ = [("z", C.c_void_p), ("mask_scale", C.c_void_p), ("mask_shift", C.c_void_p), ("mean", C.c_void_p),
                ("invstd", C.c_void_p), ("gamma", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
                ("coef", C.c_void_p), ("slots", C.c_void_p), ("ticket", C.c_void_p), ("count", C.c_int64),
                ("accumulate", C.c_int32)]


class MmrHeadMetric(C.Structure):
    _fields_ = [("labels", C.c_void_p), ("labels_u8", C.c_int32), ("pred_out", C.c_void_p),
                ("confusion", C.c_void_p)]


class MmrHaloSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
                ("N", C.c_int32), ("up", C.c_int32)]


class MmrHaloConvDesc(C.Structure):
    _fields_ = [
        ("nsrc", C.c_int32), ("src", MmrHaloSrc * 6),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("weights", C.c_void_p), ("cb", C.c_int32), ("bn", C.c_int32), ("sg", C.c_int32), ("n_ntiles", C.c_int32),
        ("tx", C.c_int32), ("tps", C.c_int32),
        ("halo_stages", C.c_int32), ("w_slots", C.c_int32), ("acc_bufs", C.c_int32),
        ("out_stages", C.c_int32), ("direct_store", C.c_int32),
        ("ngroups", C.c_int32), ("groups", C.POINTER(MmrOutSeg)),
        ("cout_total", C.c_int32),
        ("scale", C.c_void_p), ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("res_ldc", C.c_int32), ("relu", C.c_int32),
        ("out_mode", C.c_int32), ("out_f32", C.c_void_p), ("out_ldc", C.c_int32),
        ("stats", C.c_void_p), ("stats_ld", C.c_int32),
        ("bn_finalize", C.POINTER(MmrBnFinalize)),
        ("rph", C.c_int32),
        ("bn_bwd", C.POINTER(MmrBnBwdFused)),
        ("head_metric", C.POINTER(MmrHeadMetric)),
        ("loader", C.c_int32),
    ]


class MmrPackJob(C.Structure):
    _fields_ = [("w_oihw", C.c_void_p), ("out", C.c_void_p), ("O", C.c_int32), ("I", C.c_int32),
                ("mode", C.c_int32), ("cb", C.c_int32), ("bn", C.c_int32), ("n_ntiles", C.c_int32),
                ("nchunks", C.c_int32), ("layout", C.c_int32)]


class MmrBnFoldJob(C.Structure):
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("conv_bias", C.c_void_p), ("scale", C.c_void_p),
                ("shift", C.c_void_p), ("C", C.c_int32), ("rep", C.c_int32)]


class MmrWgradHaloDesc(C.Structure):
    _fields_ = [
        ("dz", MmrHaloSrc), ("nsrc", C.c_int32), ("src", MmrHaloSrc * 6),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("cb", C.c_int32), ("bn", C.c_int32), ("cout_gemm", C.c_int32), ("tx", C.c_int32),
        ("n_split", C.c_int32), ("partial", C.c_void_p), ("dst", C.c_void_p),
        ("dst_cout", C.c_int32), ("dst_cin", C.c_int32), ("mode", C.c_int32),
    ]


class MmrWgChunk(C.Structure):
    _fields_ = [("src", C.c_int32), ("c0", C.c_int32), ("a", C.c_int32), ("bx", C.c_int32 * 4),
                ("by", C.c_int32 * 4), ("dst_ci", C.c_int32), ("dst_tap", C.c_int32)]


class MmrWgradDesc(C.Structure):
    _fields_ = [
        ("dz", MmrSrc), ("dz_a", C.c_int32), ("dz_bx", C.c_int32 * 4), ("dz_by", C.c_int32 * 4),
        ("nsrc", C.c_int32), ("src", MmrSrc * 6),
        ("ncls", C.c_int32),
        ("nchunks", C.c_int32), ("chunks", C.POINTER(MmrWgChunk)),
        ("chunk_ch", C.c_int32), ("cout_gemm", C.c_int32),
        ("kp_w", C.c_int32), ("kp_h", C.c_int32), ("kp_n", C.c_int32),
        ("gx_count", C.c_int32), ("gy_count", C.c_int32), ("n_img", C.c_int32),
        ("n_split", C.c_int32), ("partial", C.c_void_p), ("dst", C.c_void_p),
        ("dst_cout", C.c_int32), ("dst_cin", C.c_int32), ("dst_taps", C.c_int32),
        ("chunk_valid_ch", C.c_int32),
    ]


class MmrContrib(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("pool2", C.c_int32)]


class MmrLossParams(C.Structure):
    _fields_ = [("dice_eps_nr", C.c_float), ("dice_eps_dr", C.c_float), ("onehot_eps", C.c_float),
                ("w_dice", C.c_float), ("w_ce", C.c_float), ("dice_channels", C.c_int32),
                ("ce_ignore_index", C.c_int64)]


MMR_OUT_BF16_NHWC = 0
MMR_OUT_F32_NCHW = 1

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol include/mmrseg.h declares.
SIGNATURES = {
    "mmr_last_error": (C.c_char_p, []),
    "mmr_abi_version": (_i, []),
    "mmr_device_ok": (_i, []),
    "mmr_conv_plan_create": (_i, [C.POINTER(MmrConvDesc), C.POINTER(_vp)]),
    "mmr_conv_plan_run": (_i, [_vp, _i, _vp]),
    "mmr_conv_plan_destroy": (_i, [_vp]),
    "mmr_halo_conv_plan_create": (_i, [C.POINTER(MmrHaloConvDesc), C.POINTER(_vp)]),
    "mmr_halo_conv_plan_run": (_i, [_vp, _vp]),
    "mmr_halo_conv_plan_destroy": (_i, [_vp]),
    "mmr_debug_halo_trace": (_i, [_vp, _i]),
    "mmr_pack_weights_halo": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "mmr_pack_items_per_block": (_i, []),
    "mmr_pack_weights_halo_batch": (_i, [_vp, _i, _i64, _vp]),
    "mmr_wgrad_halo_partial_floats": (_i64, [_i, _i, _i, _i, _i]),
    "mmr_wgrad_kx_partial_floats": (_i64, [_i, _i, _i, _i]),
    "mmr_wgrad_thin_partial_floats": (_i64, [_i, _i, _i, _i, _i]),
    "mmr_wgrad_halo_plan_create": (_i, [C.POINTER(MmrWgradHaloDesc), C.POINTER(_vp)]),
    "mmr_wgrad_halo_plan_run": (_i, [_vp, _i, _vp]),
    "mmr_wgrad_halo_plan_destroy": (_i, [_vp]),
    "mmr_wgrad_plan_create": (_i, [C.POINTER(MmrWgradDesc), C.POINTER(_vp)]),
    "mmr_wgrad_plan_run": (_i, [_vp, _i, _i, _vp]),
    "mmr_wgrad_plan_destroy": (_i, [_vp]),
    "mmr_stem_im2col": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "mmr_stem_s2d_pack": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "mmr_stem_s2d_weights": (_i, [_vp, _i, _vp, _vp]),
    "mmr_stem_s2d_wgrad_fold": (_i, [_vp, _i, _vp, _i, _vp]),
    "mmr_stem_im2col_u8": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "mmr_pack_nhwc_u8_to_nhwc_bf16": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "mmr_pack_nchw_f32_to_nhwc_bf16": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "mmr_unpack_nhwc_bf16_to_nchw_f32": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "mmr_repack_weights": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _i, _vp]),
    "mmr_bn_stats": (_i, [_vp, _i64, _i, _vp, _i, _vp]),
    "mmr_bn_finalize": (_i, [_vp, _i, _i64, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                             _vp, _vp]),
    "mmr_bn_apply": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    "mmr_bn_bwd_reduce": (_i, [C.POINTER(MmrContrib), _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp,
                               _vp, _i, _vp]),
    "mmr_bn_bwd_reduce_fused": (_i, [C.POINTER(MmrContrib), _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i,
                                     _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "mmr_bn_bwd_finalize": (_i, [_vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "mmr_bn_bwd_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp]),
    "mmr_bn_bwd_apply_masked": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp]),
    "mmr_grad_gather": (_i, [C.POINTER(MmrContrib), _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "mmr_bias_grad_finalize": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "mmr_maxpool3x3s2_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "mmr_maxpool3x3s2_bwd": (_i, [C.POINTER(MmrContrib), _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "mmr_window_gather": (_i, [_vp, _i, _i, _i, _i, C.POINTER(C.c_int), _i, C.POINTER(C.c_int), _i, _i, _i, _i, _i,
                               _vp, _vp]),
    "mmr_window_blend": (_i, [_vp, _i, _i, _i, _i, C.POINTER(C.c_int), _i, C.POINTER(C.c_int), _i, _i, _i, _vp, _vp,
                              _vp]),
    "mmr_maxpool2x2s2_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "mmr_maxpool2x2s2_bwd": (_i, [C.POINTER(MmrContrib), _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "mmr_upsample_bilinear2x_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "mmr_upsample_bilinear2x_bwd": (_i, [C.POINTER(MmrContrib), _i, _i, _i, _i, _i, _vp, _vp]),
    "mmr_upsample_nearest_f32_nchw": (_i, [_vp, _i64, _i, _i, _i, _vp, _vp]),
    "mmr_sumpool_f32_nchw": (_i, [_vp, _i64, _i, _i, _i, _vp, _vp]),
    "mmr_dice_ce_workspace_doubles": (_i64, [_i, _i, _i]),
    "mmr_dice_ce_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, C.POINTER(MmrLossParams), _vp, _i, _vp,
                             _vp]),
    "mmr_dice_ce_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, C.POINTER(MmrLossParams), _vp, _f, _vp,
                             _vp, _vp]),
    "mmr_head_grad_prep": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp]),
    "mmr_pointwise_head_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "mmr_pointwise_head_bwd_workspace_bytes": (_i64, [_i]),
    "mmr_pointwise_head_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "mmr_confusion_from_logits": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "mmr_confusion_from_preds": (_i, [_vp, _vp, _i, _i, _i64, _i64, _i, _vp, _vp]),
    "mmr_onehot_to_labels": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "mmr_hausdorff_workspace_bytes": (_i64, [_i, _i, _i]),
    "mmr_hausdorff_sq": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i64, _vp, _vp]),
    "mmr_adam_step": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _f, _f, _i, _f, _vp]),
    "mmr_zero_async": (_i, [_vp, _i64, _vp]),
    "mmr_sgd_step": (_i, [_vp, _vp, _vp, _i64, _f, _f, _f, _i, _f, _vp]),
    "mmr_sumsq": (_i, [_vp, _i64, _vp, _vp]),
    "mmr_clip_scale": (_i, [_vp, _i64, _vp, _f, _vp, _vp]),
    "mmr_bn_fold_batch": (_i, [_vp, _i, _f, _vp]),
}

_lib = None


def lib():
    """Load libmmrseg.so once.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmrError(
                "libmmrseg.so not found at %s: build it with `python __graft_entry__.py` "
                "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise MmrError(lib().mmr_last_error().decode(errors="replace"))


def device_ok():
    return bool(lib().mmr_device_ok())
