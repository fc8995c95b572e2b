import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mmrseg_b200.data import DevicePrefetcher
from mmrseg_b200.losses import DiceCrossEntropyLoss
from mmrseg_b200.models import UnetPlusPlus
from mmrseg_b200.optim import FusedAdam

dev = torch.device("cuda", 0)
n = bench.BATCH_PER_GPU
model = UnetPlusPlus("resnet18", classes=2).to(dev).train()
crit = DiceCrossEntropyLoss(0.5)
opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
xh, yh = bench.synthetic(n, 6210, pinned=True)


def step(x, y):
    for p in model.parameters():
        p.grad = None
    loss = crit(model(x), y)
    loss.backward()
    opt.step()
    return loss


xd, yd = xh.to(dev), yh.to(dev)
for _ in range(4):
    step(xd, yd).item()
mode = sys.argv[1]
pf = DevicePrefetcher([(xh, yh)] * 12, dev)
if mode == "mainstream":
    pf.stream = torch.cuda.current_stream()
t_prev = time.perf_counter()
for i, (x, y) in enumerate(pf):
    t0 = time.perf_counter()
    l = step(x, y)
    t1 = time.perf_counter()
    l.item()
    t2 = time.perf_counter()
    print("%s iter %2d: fetch %.2f ms, issue %.2f ms, wait %.2f ms" % (mode, i, (t0 - t_prev) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    t_prev = t2
