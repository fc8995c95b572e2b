"""The other BASELINE configs at their full sizes on one GPU (parity cases for the tests, measured here for
DESIGN.md): C3 ResNetUNet-34 train step batch 32 @ 512^2 (10 classes), C4's per-GPU share (U-Net++ with deep
supervision, 32 images = global 256 over 8 GPUs), the in-tree UNet train step batch 16 @ 512^2, and -- through
scripts/infer_bench.py -- C5.  usage: python scripts/config_bench.py [c3|c4|unet ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmrseg_b200.losses import DiceCrossEntropyLoss
from mmrseg_b200.models import ResNetUNet, UNet, UnetPlusPlus
from mmrseg_b200.optim import FusedAdam


def run(name, model, n, classes, hw=512, steps=8):
    g = torch.Generator().manual_seed(6210)
    x = torch.randn((n, 3, hw, hw), generator=g).cuda()
    y = torch.randint(0, classes, (n, hw, hw), generator=g).cuda()
    model = model.cuda().train()
    crit = DiceCrossEntropyLoss(0.5)
    opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)

    def step():
        for p in model.parameters():
            p.grad = None
        out = model(x)
        outs = out if isinstance(out, list) else [out]
        loss = sum(crit(o, y) for o in outs)
        loss.backward()
        opt.step()
        return loss

    for _ in range(4):
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng = list(model._engines.values())[0]
    fl = eng.conv_flops_fwd + eng.conv_flops_bwd
    print("%-44s batch %3d @ %d: %8.2f ms/step %8.1f img/s  %6.1f conv TFLOP/s  loss %.4f  mem %.1f GB" % (
        name, n, hw, ms, n / ms * 1e3, fl / ms / 1e9, float(loss), torch.cuda.max_memory_allocated() / 1e9), flush=True)
    del model, opt
    torch.cuda.empty_cache()


which = sys.argv[1:] or ["c3", "c4", "unet"]
torch.manual_seed(6210)
if "c3" in which:
    run("C3 ResNetUNet-34 train (10 classes)", ResNetUNet(10, 34), 32, 10)
if "c4" in which:
    run("C4 U-Net++ deep supervision train, 1/8 share", UnetPlusPlus("resnet18", classes=2, deep_supervision=True), 32, 2)
if "unet" in which:
    run("in-tree UNet train (10 classes)", UNet(3, 10, bilinear=True), 16, 10)
