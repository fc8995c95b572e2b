"""CUDA-event time of one train step per C-ABI entry point.

    python scripts/step_breakdown.py [batch] [config]      config = c2 (default) | c3 | c4"""
import ctypes as C
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
CFG = bench.resolve(sys.argv[2] if len(sys.argv) > 2 else "c2", 1)
from mmrseg_b200.losses import DiceCrossEntropyLoss
from mmrseg_b200.optim import FusedAdam

n = int(sys.argv[1]) if len(sys.argv) > 1 and int(sys.argv[1]) > 0 else CFG["batch"]
torch.manual_seed(6210)
model = bench.build_model(CFG, torch.device("cuda", 0)).train()
crit = DiceCrossEntropyLoss(0.5)
opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
x, y = bench.synthetic(CFG, n)
x, y = x.cuda(), y.cuda()
for i in range(3):
    for p in model.parameters():
        p.grad = None
    out = model(x)
    loss = sum(crit(o, y) for o in out) / len(out) if isinstance(out, list) else crit(out, y)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5):
    for p in model.parameters():
        p.grad = None
    out = model(x)
    loss = sum(crit(o, y) for o in out) / len(out) if isinstance(out, list) else crit(out, y)
    loss.backward()
    opt.step()
e1.record()
torch.cuda.synchronize()
print("step %.3f ms (%.1f img/s)" % (e0.elapsed_time(e1) / 5, n * 5 / e0.elapsed_time(e1) * 1e3))
eng = list(model._engines.values())[0]
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
agg = defaultdict(lambda: [0.0, 0])
for it in range(3):
    evs = []
    for phase, calls in (("repack", eng.repack_calls), ("fwd", eng.fwd_calls), ("bwd", eng.bwd_calls[False])):
        for fn, a in calls:
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            fn(*a, sp)
            a1.record(stream)
            evs.append((phase + ":" + fn.__name__, a0, a1))
    torch.cuda.synchronize()
    if it > 0:
        for name, a0, a1 in evs:
            agg[name][0] += a0.elapsed_time(a1) / 2
            agg[name][1] += 0.5
tot = sum(v[0] for v in agg.values())
for name, (ms, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-40s %8.3f ms  %5.1f %%  %4d launches" % (name, ms, 100 * ms / tot, cnt))
print("sum of per-call times %.3f ms" % tot)

# the same launch list replayed from a CUDA graph: how much of the step is launch gaps
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    sp2 = C.c_void_p(side.cuda_stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for calls in (eng.repack_calls, eng.fwd_calls, eng.bwd_calls[False]):
            for fn, a in calls:
                fn(*a, sp2)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
for label, fn in (("graph replay", graph.replay),):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%s of repack+fwd+bwd: %.3f ms" % (label, e0.elapsed_time(e1) / 10))


def eager():
    for calls in (eng.repack_calls, eng.fwd_calls, eng.bwd_calls[False]):
        for fn, a in calls:
            fn(*a, sp)


for _ in range(3):
    eager()
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    eager()
e1.record()
torch.cuda.synchronize()
print("eager launch list of repack+fwd+bwd: %.3f ms" % (e0.elapsed_time(e1) / 10))
