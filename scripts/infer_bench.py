"""BASELINE configs[4]: U-Net++ multiclass inference throughput at 1024x1280, batch 64, BN folded,
argmax + confusion-matrix metric on the device.  usage: python scripts/infer_bench.py [batch] [H] [W]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmrseg_b200.metrics import confusion_matrix
from mmrseg_b200.models import UnetPlusPlus

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1280
classes = 10
torch.manual_seed(6210)
model = UnetPlusPlus("resnet18", classes=classes).cuda().eval()
g = torch.Generator().manual_seed(6210)
x = torch.rand((n, 3, H, W), generator=g).cuda()
y = torch.randint(0, classes, (n, H, W), generator=g).cuda()
cm = torch.zeros((n, classes, classes), device="cuda", dtype=torch.int64)
with torch.no_grad():
    for _ in range(2):
        logits = model(x)
        confusion_matrix(logits, y, cm=cm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    e0.record()
    for _ in range(iters):
        logits = model(x)
        confusion_matrix(logits, y, cm=cm)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
eng = list(model._engines.values())[0]
fl = eng.conv_flops_fwd
print("inference %dx%dx%d: %.2f ms/batch, %.1f img/s, %.1f TFLOP/s conv, mem %.1f GB, cm sum %d" % (
    n, H, W, ms, n / ms * 1e3, fl / ms / 1e9, torch.cuda.max_memory_allocated() / 1e9, int(cm.sum())))
