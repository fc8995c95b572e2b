"""Per-layer table of the tcgen05 conv kernels at the bench configuration: CUDA-event time of every
fprop / dgrad / wgrad launch (L2 flushed between launches), algorithmic TFLOP/s and fraction of
the measured bf16 peak.  Writes gpurun_out/layers_<tag>.csv.

    python scripts/layer_table.py [tag] [batch] [config]      config = c2 (default) | c3 | c4"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
cfg = bench.resolve(sys.argv[3] if len(sys.argv) > 3 else "c2", 1)
n = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else cfg["batch"]
reps = 3
torch.manual_seed(6210)
model = bench.build_model(cfg, torch.device("cuda", 0)).train()
x, y = bench.synthetic(cfg, n)
x = x.cuda()
eng = model._engine_for(x, training=True)
eng.forward(x)
eng.backward(torch.zeros_like(eng.acts["logits"].buf)) if not cfg["ds"] else None
torch.cuda.synchronize()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
pk = bench.peaks()["tflops"]


def time_call(plan):
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        plan.run(stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


rows = []
lib = eng.lib
for u in eng.units:
    if "fplan" not in u:
        continue
    name = u["op"]["conv"]
    for kind, plan in (("fprop", u.get("fplan")), ("dgrad", u.get("dplan")), ("wgrad", u.get("wplan"))):
        if plan is None:
            continue
        ms = time_call(plan)
        tf = plan.flops / (ms * 1e-3) / 1e12
        rows.append((name, kind, plan.flops / 1e9, ms, tf, tf / pk))
os.makedirs("profiles", exist_ok=True)
out = os.path.join("gpurun_out", "layers_%s.csv" % tag)
with open(out, "w") as f:
    f.write("layer,kind,gflop,ms,tflops,frac_of_measured_sustained_peak\n")
    for r in rows:
        f.write("%s,%s,%.2f,%.4f,%.1f,%.3f\n" % r)
tot = {}
for name, kind, gf, ms, tf, fr in rows:
    t = tot.setdefault(kind, [0.0, 0.0])
    t[0] += gf
    t[1] += ms
for kind, (gf, ms) in tot.items():
    print("%s: %.1f GFLOP in %.3f ms = %.1f TFLOP/s (%.3f of %.1f)" % (kind, gf, ms, gf / ms, gf / ms / pk, pk))
print("all: %.3f ms" % sum(v[1] for v in tot.values()))
for r in sorted(rows, key=lambda r: -r[3])[:40]:
    print("%-45s %-6s %8.1f GF %8.3f ms %7.1f TF/s %.3f" % r)
