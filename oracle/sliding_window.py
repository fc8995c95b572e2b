"""Oracle restatement of monai.inferers.sliding_window_inference, mode="constant" (TEST INFRASTRUCTURE).

monai is not vendored under /root/reference and not installed (pyproject.toml:12 lists it unpinned), so this
restates the published algorithm and is anchored on the reference's call site
ED/Main_MMR_SegModel.py:1308-1320 (roi 512x640, overlap 0.5, 1080x1920 frames -> 4x5 windows per frame;
1024x1280 -> 3x3): **parity unpinned**.
"""
import math

import torch


def scan_starts(length, roi, overlap):
    interval = roi if roi == length else max(int(roi * (1 - overlap)), 1)
    num = int(math.ceil(float(length - roi) / interval)) + 1 if roi < length else 1
    return [min(i * interval, length - roi) for i in range(num)]


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25):
    n, _, h, w = inputs.shape
    rh, rw = roi_size
    slices = [(i, y, x) for i in range(n) for y in scan_starts(h, rh, overlap) for x in scan_starts(w, rw, overlap)]
    out = cnt = None
    for b in range(0, len(slices), sw_batch_size):
        chunk = slices[b:b + sw_batch_size]
        batch = torch.cat([inputs[i:i + 1, :, y:y + rh, x:x + rw] for i, y, x in chunk], 0)
        pred = predictor(batch)
        if out is None:
            out = torch.zeros((n, pred.shape[1], h, w), dtype=pred.dtype)
            cnt = torch.zeros((n, 1, h, w), dtype=pred.dtype)
        for k, (i, y, x) in enumerate(chunk):
            out[i, :, y:y + rh, x:x + rw] += pred[k]      # importance map of ones
            cnt[i, :, y:y + rh, x:x + rw] += 1.0
    return out / cnt
