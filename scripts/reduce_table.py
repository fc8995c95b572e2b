"""Every BatchNorm-backward reduction launch of one U-Net++ train step (batch 16 @ 512x512): shape, number of
gradient contributions (p = 2x2-pooled), whether g is written, CUDA-event time and algorithmic TB/s."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
CFG = bench.resolve("c2", 1)
from mmrseg_b200.losses import DiceCrossEntropyLoss
from mmrseg_b200.models import UnetPlusPlus

n = int(sys.argv[1]) if len(sys.argv) > 1 else CFG["batch"]
torch.manual_seed(6210)
model = UnetPlusPlus("resnet18", classes=CFG["classes"]).cuda().train()
crit = DiceCrossEntropyLoss(0.5)
x, y = bench.synthetic(CFG, n)
x, y = x.cuda(), y.cuda()
for i in range(2):
    for p in model.parameters():
        p.grad = None
    crit(model(x), y).backward()
torch.cuda.synchronize()
eng = list(model._engines.values())[0]
lib = eng.lib
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
rows = {}
for it in range(3):
    evs = []
    for k, (fn, a) in enumerate(eng.bwd_calls[False]):
        if fn is lib.mmr_bn_bwd_reduce_fused:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn(*a, sp)
            e1.record(stream)
            evs.append((k, a, e0, e1))
        else:
            fn(*a, sp)
    torch.cuda.synchronize()
    if it:
        for k, a, e0, e1 in evs:
            rows.setdefault(k, [a, 0.0])[1] += e0.elapsed_time(e1) / 2
tot = 0.0
for k, (a, ms) in rows.items():
    arr, cnt, act, z, mean, invstd, nn, ho, wo, cc, g = a[:11]
    pools = [arr[i].pool2 for i in range(cnt)]
    el = nn * ho * wo * cc * 2
    byts = el + sum(el * (4 if p else 1) for p in pools) + (el if act and getattr(act, "value", act) else 0) + \
        (el if g and getattr(g, "value", g) else 0)
    tot += ms
    print("%4d  %3dx%-3d C %3d  contribs %-12s act %d  g %d  %7.1f us  %5.2f TB/s" % (
        k, ho, wo, cc, "".join("p" if p else "s" for p in pools), bool(act and getattr(act, "value", act)),
        bool(g and getattr(g, "value", g)), ms * 1e3, byts / ms / 1e9))
print("total %.3f ms in %d launches" % (tot, len(rows)))
