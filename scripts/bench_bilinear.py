"""Bilinear x2 (align_corners=True) forward / adjoint kernels in isolation at the five shapes of config 3
(ResNetUNet-34, batch 32 @ 512x512): time per launch (L2 flushed between launches) and algorithmic TB/s
(forward: source read once + output written once; adjoint: gradient read once + source gradient written once)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmrseg_b200 import _lib
from mmrseg_b200._lib import MmrContrib

lib = _lib.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
shapes = [(16, 16, 512), (32, 32, 512), (64, 64, 256), (128, 128, 256), (256, 256, 128)]
if len(sys.argv) > 2:      # one shape only (for an ncu capture)
    shapes = [shapes[int(sys.argv[2])]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
tot = [0.0, 0.0]
for H, W, Cc in shapes:
    x = torch.randn((N, H, W, Cc), device="cuda").to(torch.bfloat16)
    out = torch.empty((N, 2 * H, 2 * W, Cc), device="cuda", dtype=torch.bfloat16)
    go = torch.randn((N, 2 * H, 2 * W, Cc), device="cuda").to(torch.bfloat16)
    gin = torch.empty_like(x)
    arr = (MmrContrib * 1)()
    arr[0].ptr, arr[0].pool2 = go.data_ptr(), 0
    fns = [lambda: lib.mmr_upsample_bilinear2x_fwd(C.c_void_p(x.data_ptr()), N, H, W, Cc, C.c_void_p(out.data_ptr()), s),
           lambda: lib.mmr_upsample_bilinear2x_bwd(arr, 1, N, H, W, Cc, C.c_void_p(gin.data_ptr()), s)]
    nbytes = 5 * x.numel() * 2
    res = []
    for k, fn in enumerate(fns):
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(fn())
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[1:])[len(ts[1:]) // 2]
        tot[k] += t
        res.append("%8.3f ms %5.2f TB/s" % (t, nbytes / t / 1e9))
    print("%4d x %4d x %4d   fwd %s   bwd %s" % (H, W, Cc, res[0], res[1]))
print("total fwd %.3f ms, bwd %.3f ms" % tuple(tot))
