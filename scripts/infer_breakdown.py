"""Per-launch CUDA-event times of the C5 inference plan (U-Net++ eval, batch 64 @ 1024x1280, 10 classes)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmrseg_b200.models import UnetPlusPlus

n, H, W = int(os.environ.get("BATCH", "64")), 1024, 1280
model = UnetPlusPlus("resnet18", classes=10).cuda().eval()
x = torch.rand((n, 3, H, W)).cuda()
with torch.no_grad():
    for _ in range(2):
        model(x)
eng = list(model._engines.values())[0]
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
names = {}
for k in dir(eng.lib):
    pass
rows = []
units = [u for u in eng.units if "fplan" in u]
ui = 0
for fn, a in eng.fwd_calls:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn(*a, sp)
    e1.record(stream)
    torch.cuda.synchronize()
    name = getattr(fn, "__name__", str(fn))
    label = name
    if "conv_plan_run" in name and ui < len(units):
        u = units[ui]
        ui += 1
        label = "%s %s %.0f GF" % (name.replace("mmr_", "").replace("_plan_run", ""), u["op"]["conv"], u["fplan"].flops / 1e9)
        rows.append((e0.elapsed_time(e1), label, u["fplan"].flops))
    else:
        rows.append((e0.elapsed_time(e1), label, 0))
tot = sum(r[0] for r in rows)
print("total %.2f ms over %d launches" % (tot, len(rows)))
for ms, label, fl in sorted(rows, reverse=True)[:16]:
    print("%8.3f ms %5.1f %%  %-60s %s" % (ms, 100 * ms / tot, label, ("%.0f TF/s" % (fl / ms / 1e9)) if fl else ""))
