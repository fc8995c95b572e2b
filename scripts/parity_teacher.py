"""Print the teacher-forced layer-local parity table (tests/teacher.py) of one model / shape: the worst error per
(quantity) over all units, and every row above a threshold.  usage:
  MMR_NO_ARENA_REUSE=1 python scripts/parity_teacher.py [unetpp18|unetpp34|unetpp18ds|resnet_unet34|unet] N H W [C]"""
import os
import sys

os.environ["MMR_NO_ARENA_REUSE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tests import teacher  # noqa: E402
from tests.helpers import model_pair, synthetic_batch  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "unetpp18"
n, h, w = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (2, 128, 128)
classes = int(sys.argv[5]) if len(sys.argv) > 5 else 2
if which.startswith("unetpp"):
    enc = "resnet34" if "34" in which else "resnet18"
    if which.endswith("ds"):
        from oracle.unetpp import DeepSupervisionUnetPlusPlus
        from mmrseg_b200.models import UnetPlusPlus
        torch.manual_seed(6210)
        ref = DeepSupervisionUnetPlusPlus(enc, None, 3, classes)
        net = UnetPlusPlus(enc, classes=classes, deep_supervision=True)
        net.load_state_dict(ref.state_dict(), strict=True)
        net = net.cuda()
    else:
        _, net = model_pair(classes, enc)
elif which == "resnet_unet34":
    from tests.test_resnet_unet_gpu import _pair
    net = _pair(classes, 34)[1].cuda()
else:
    from tests.test_unet_gpu import _pair
    net = _pair(classes)[1].cuda()
from mmrseg_b200.losses import DiceCrossEntropyLoss  # noqa: E402
x, y = synthetic_batch(n, classes, h, w)
net.train()
P = teacher.snapshot(net)
out = net(x.cuda())
outs = out if isinstance(out, list) else [out]
crit = DiceCrossEntropyLoss(0.5)
(sum(crit(o, y.cuda()) for o in outs) / len(outs)).backward()
torch.cuda.synchronize()
eng = [e for k, e in net._engines.items() if k[3]][0]
rows = teacher.check_forward(eng, P, x) + teacher.check_backward(eng, P, teacher.grads_of(net), x)
by_what = {}
for unit, what, kind, err in rows:
    key = (kind, what.split(" -> ")[0])
    cur = by_what.setdefault(key, [0, 0.0, None, []])
    cur[0] += 1
    cur[3].append(err)
    if err >= cur[1]:
        cur[1], cur[2] = err, unit
print("%s %dx%dx%d C=%d: %d comparisons over %d units" % (which, n, h, w, classes, len(rows), len(eng.units)))
print("%-6s %-46s %5s %10s %10s  %s" % ("kind", "quantity", "n", "median", "worst", "worst unit"))
for (kind, what), (cnt, worst, unit, errs) in sorted(by_what.items()):
    errs.sort()
    print("%-6s %-46s %5d %10.2e %10.2e  %s" % (kind, what, cnt, errs[len(errs) // 2], worst, unit))
thr = float(os.environ.get("THR", "5e-4"))
for unit, what, kind, err in rows:
    if err > thr:
        print("  > %.0e: %-30s %-34s %-5s %.3e" % (thr, unit, what, kind, err))
