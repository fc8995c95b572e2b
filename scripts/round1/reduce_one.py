"""One BatchNorm-backward reduction (z mask, no g, one same-resolution contribution, 16 x 256 x 256 x 64) a few times:
the command ncu wraps to look at that kernel alone."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mmrseg_b200 import _lib
from mmrseg_b200._lib import MmrContrib

lib = _lib.lib()
N, H, Cc = 16, 256, 64
vp = lambda t: C.c_void_p(t.data_ptr())
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
z = torch.randn((N, H, H, Cc), device="cuda").to(torch.bfloat16)
dx = torch.randn((N, H, H, Cc), device="cuda").to(torch.bfloat16)
st = torch.rand((7, Cc), device="cuda") + 0.5
arr = (MmrContrib * 1)()
arr[0].ptr, arr[0].pool2 = dx.data_ptr(), 0
slots = torch.zeros((8 * 2 * Cc,), device="cuda", dtype=torch.float64)
ticket = torch.zeros((1,), device="cuda", dtype=torch.int32)
dgb = torch.zeros((2, Cc), device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    lib.mmr_bn_bwd_reduce_fused(arr, 1, None, vp(z), vp(st[0]), vp(st[1]), N, H, H, Cc, None, vp(slots), 296,
                                vp(st[2]), vp(dgb[0]), vp(dgb[1]), 0, vp(st[4]), vp(ticket), vp(st[2]), vp(st[3]), s)
torch.cuda.synchronize()
print("ok")
