"""cProfile of the host side of the train step (what runs between two GPU graph launches)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from mmrseg_b200.losses import DiceCrossEntropyLoss
from mmrseg_b200.models import UnetPlusPlus
from mmrseg_b200.optim import FusedAdam

model = UnetPlusPlus("resnet18", classes=2).cuda().train()
crit = DiceCrossEntropyLoss(0.5)
opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
x, y = bench.synthetic(16)
x, y = x.cuda(), y.cuda()


def step():
    for p in model.parameters():
        p.grad = None
    loss = crit(model(x), y)
    loss.backward()
    opt.step()
    return loss


for _ in range(5):
    step().item()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step().item()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
