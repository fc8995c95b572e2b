"""One process, a few train steps at a bench configuration (default c2: U-Net++, batch 16 @ 512x512): the
command that the ncu launch list / full capture under profiles/ are taken from.

    python scripts/profile_step.py [batch] [steps] [config]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
CFG = bench.resolve(sys.argv[3] if len(sys.argv) > 3 else "c2", 1)
from mmrseg_b200.losses import DiceCrossEntropyLoss
from mmrseg_b200.optim import FusedAdam

n = int(sys.argv[1]) if len(sys.argv) > 1 and int(sys.argv[1]) > 0 else CFG["batch"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(6210)
model = bench.build_model(CFG, torch.device("cuda", 0)).train()
crit = DiceCrossEntropyLoss(0.5)
opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
x, y = bench.synthetic(CFG, n)
x, y = x.cuda(), y.cuda()
for i in range(steps):
    for p in model.parameters():
        p.grad = None
    out = model(x)
    loss = sum(crit(o, y) for o in out) / len(out) if isinstance(out, list) else crit(out, y)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
eng = list(model._engines.values())[0]
print("loss", float(loss), "launch calls fwd/bwd", len(eng.fwd_calls), len(eng.bwd_calls[False]))
