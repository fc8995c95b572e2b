"""Where the end-to-end step goes: resident inputs vs blocking copies vs DevicePrefetcher, fp32 and uint8 frames."""
import os, sys
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
model.set_input_normalization((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
crit = DiceCrossEntropyLoss(0.5)
opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
xh, yh = bench.synthetic(n, 6210, pinned=True)
fh = (torch.rand((n, 512, 512, 3)) * 255).to(torch.uint8).pin_memory()


def step(x, y):
    for p in model.parameters():
        p.grad = None
    loss = crit(model(x), y)
    loss.backward()
    opt.step()
    return loss


def timeit(fn, steps=10):
    fn(4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(steps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for name, hx in (("fp32", xh), ("uint8", fh)):
    xd, yd = hx.to(dev), yh.to(dev)
    print(name, "resident           %.3f ms" % timeit(lambda k: [step(xd, yd) for _ in range(k)]))
    print(name, "resident + item    %.3f ms" % timeit(lambda k: [step(xd, yd).item() for _ in range(k)]))
    print(name, "blocking copies    %.3f ms" % timeit(
        lambda k: [step(hx.to(dev, non_blocking=True), yh.to(dev, non_blocking=True)).item() for _ in range(k)]))
    print(name, "prefetcher         %.3f ms" % timeit(
        lambda k: [step(x, y).item() for x, y in DevicePrefetcher([(hx, yh)] * k, dev)]))
    print(name, "copy alone         %.3f ms" % timeit(
        lambda k: [(hx.to(dev, non_blocking=True), yh.to(dev, non_blocking=True)) for _ in range(k)]))
