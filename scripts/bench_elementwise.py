"""Bandwidth of the BatchNorm forward / backward kernels on U-Net++ activation shapes (batch 16)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmrseg_b200 import _lib
from mmrseg_b200._lib import MmrContrib

lib = _lib.lib()
N = 16
# (H, C, [pool2 flags of the contributions])
SHAPES = [(512, 16, [0]), (256, 32, [1]), (256, 64, [0]), (256, 64, [0, 1]), (128, 64, [0, 0, 0, 1]),
          (128, 64, [0]), (64, 128, [0, 0, 1]), (32, 256, [0, 1]), (16, 512, [0, 0])]


def vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
partial = torch.empty((592 * 2 * 512,), device="cuda", dtype=torch.float64)
for H, Cc, pools in SHAPES:
    P = N * H * H
    z = torch.randn((N, H, H, Cc), device="cuda").to(torch.bfloat16)
    a = torch.relu(z)
    g = torch.empty_like(z)
    dz = torch.empty_like(z)
    st = torch.rand((7, Cc), device="cuda") + 0.5
    contribs = [torch.randn((N, H * (2 if p else 1), H * (2 if p else 1), Cc), device="cuda").to(torch.bfloat16)
                for p in pools]
    arr = (MmrContrib * len(pools))()
    for i, (t, p) in enumerate(zip(contribs, pools)):
        arr[i].ptr, arr[i].pool2 = t.data_ptr(), p
    rows_per_iter = 256 // (Cc // 8)
    nblk = int(max(1, min(296, -(-P // (rows_per_iter * 4)))))
    el = P * Cc * 2
    cb = sum(t.numel() * 2 for t in contribs)
    t_red = timeit(lambda: lib.mmr_bn_bwd_reduce(arr, len(pools), vp(a), vp(z), vp(st[0]), vp(st[1]), N, H, H, Cc,
                                                 vp(g), vp(partial), nblk, s))
    t_bap = timeit(lambda: lib.mmr_bn_bwd_apply(vp(g), vp(z), vp(st[0]), vp(st[1]), vp(st[4]), P, Cc, vp(dz), s))
    t_bam = timeit(lambda: lib.mmr_bn_bwd_apply_masked(vp(contribs[0] if not pools[0] else g), vp(z), vp(st[0]), vp(st[1]),
                                                       vp(st[4]), vp(st[2]), vp(st[3]), P, Cc, vp(dz), s))
    slots = torch.zeros((8 * 2 * Cc,), device="cuda", dtype=torch.float64)
    ticket = torch.zeros((1,), device="cuda", dtype=torch.int32)
    dgb = torch.zeros((2, Cc), device="cuda")
    fused = lambda gout, masked: lib.mmr_bn_bwd_reduce_fused(
        arr, len(pools), None if masked else vp(a), vp(z), vp(st[0]), vp(st[1]), N, H, H, Cc, gout, vp(slots), nblk,
        vp(st[2]), vp(dgb[0]), vp(dgb[1]), 0, vp(st[4]), vp(ticket), vp(st[2]) if masked else None,
        vp(st[3]) if masked else None, s)
    t_f1 = timeit(lambda: fused(vp(g), False))
    t_f3 = timeit(lambda: fused(vp(g), True))
    t_f3n = timeit(lambda: fused(None, True))
    print("   fused reduce: act mask + g %6.1f us %5.2f TB/s | z mask + g %6.1f us %5.2f TB/s | z mask, no g %6.1f us %5.2f TB/s"
          % (t_f1 * 1e3, (cb + 3 * el) / t_f1 / 1e9, t_f3 * 1e3, (cb + 2 * el) / t_f3 / 1e9, t_f3n * 1e3,
             (cb + el) / t_f3n / 1e9), flush=True)
    t_app = timeit(lambda: lib.mmr_bn_apply(vp(z), P, Cc, vp(st[2]), vp(st[3]), None, 1, vp(a), s))
    t_sta = timeit(lambda: lib.mmr_bn_stats(vp(z), P, Cc, vp(partial), nblk, s))
    print("H %4d C %4d pools %-14s reduce %7.1f us %5.2f TB/s | bwd_apply %6.1f us %5.2f TB/s | apply %6.1f us %5.2f TB/s"
          " | stats %6.1f us %5.2f TB/s | bwd_apply_masked %6.1f us %5.2f TB/s" % (
              H, Cc, pools, t_red * 1e3, (cb + 3 * el) / t_red / 1e9, t_bap * 1e3,
              3 * el / t_bap / 1e9, t_app * 1e3, 2 * el / t_app / 1e9, t_sta * 1e3,
              el / t_sta / 1e9, t_bam * 1e3, 3 * el / t_bam / 1e9), flush=True)
