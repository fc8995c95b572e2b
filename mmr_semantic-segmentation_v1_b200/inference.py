"""Sliding-window inference with monai's signature, on the device.

    preds = sliding_window_inference(frames, roi_size=patch_size, sw_batch_size=..., predictor=model,
                                     overlap=sw_overlap)                 ED/Main_MMR_SegModel.py:1308-1317
    preds = preds.argmax(1)                                              ED/Main_MMR_SegModel.py:1320

Window placement follows monai (`dense_patch_slices` / `_get_scan_interval`): per axis the scan interval is
int(roi * (1 - overlap)) (at least 1), ceil((L - roi) / interval) + 1 starts at multiples of the interval, the
last ones clamped to L - roi.  All windows of all frames are flattened frame-major and pushed through the
predictor `sw_batch_size` at a time; the constant importance map makes the result the plain mean of the window
logits covering each pixel.  The gather, the blend and the argmax are kernels of csrc/sliding_window.cu; the
blend is a per-pixel sum in window order (deterministic).  Frames smaller than the window (monai pads them) are
not built: the reference's frames are 1080x1920 against 512x640 windows.
"""
import ctypes as C
import math

import torch

from . import _lib
from .losses import _stream


def scan_starts(length, roi, overlap):
    """Window starts along one axis (monai `_get_scan_interval` + `dense_patch_slices`)."""
    if roi > length:
        raise _lib.MmrError("sliding window of %d does not fit an axis of %d (padding small frames is not built)"
                            % (roi, length))
    interval = roi if roi == length else max(int(roi * (1 - overlap)), 1)
    num = int(math.ceil(float(length - roi) / interval)) + 1 if roi < length else 1
    return [min(i * interval, length - roi) for i in range(num)]


def _ints(v):
    return (C.c_int * len(v))(*v)


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25, mode="constant",
                             return_argmax=False, **unused):
    """inputs: fp32 [N,3,H,W] (or uint8 [N,H,W,3] frames) on the device; predictor: a model of this package in
    eval mode (any callable returning fp32 [B,C,rh,rw] logits works).  Returns the blended logits [N,C,H,W]
    (monai's return value); with return_argmax=True returns (logits, argmax int64 [N,H,W]) from the same kernel."""
    if mode != "constant":
        raise NotImplementedError("only mode='constant' (the reference's default) is built")
    if not inputs.is_cuda:
        raise _lib.MmrError("sliding_window_inference runs on a B200 only (input on %s); there is no CPU fallback"
                            % inputs.device)
    u8 = inputs.dtype == torch.uint8
    inputs = inputs.contiguous() if u8 else inputs.contiguous().float()
    if u8:
        n, h, w, _ = inputs.shape
    else:
        n, _, h, w = inputs.shape
    rh, rw = (roi_size, roi_size) if isinstance(roi_size, int) else tuple(roi_size)
    ys, xs = scan_starts(h, rh, overlap), scan_starts(w, rw, overlap)
    if len(ys) > 16 or len(xs) > 16:
        raise _lib.MmrError("more than 16 window starts per axis")
    total = n * len(ys) * len(xs)
    lib = _lib.lib()
    cys, cxs = _ints(ys), _ints(xs)
    batch = torch.empty((sw_batch_size, rh, rw, 3) if u8 else (sw_batch_size, 3, rh, rw), device=inputs.device,
                        dtype=inputs.dtype)
    win = None
    for w0 in range(0, total, sw_batch_size):
        _lib.check(lib.mmr_window_gather(inputs.data_ptr(), int(u8), n, h, w, cys, len(ys), cxs, len(xs), rh, rw,
                                         w0, sw_batch_size, batch.data_ptr(), _stream()))
        with torch.no_grad():
            logits = predictor(batch)
        if win is None:
            c = logits.shape[1]
            win = torch.empty((total, c, rh, rw), device=inputs.device, dtype=torch.float32)
        k = min(sw_batch_size, total - w0)
        win[w0:w0 + k].copy_(logits[:k])
    out = torch.empty((n, c, h, w), device=inputs.device, dtype=torch.float32)
    pred = torch.empty((n, h, w), device=inputs.device, dtype=torch.int64) if return_argmax else None
    _lib.check(lib.mmr_window_blend(win.data_ptr(), n, c, h, w, cys, len(ys), cxs, len(xs), rh, rw, out.data_ptr(),
                                    pred.data_ptr() if pred is not None else None, _stream()))
    return (out, pred) if return_argmax else out
