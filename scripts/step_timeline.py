"""Kernel timeline of graph-replayed train steps (CUPTI activity records through torch.profiler; nsys is not in the
image): for every kernel its stream, start and duration, and from those -- per step -- the busy time of each
stream, the gaps between consecutive kernels of the main stream (who follows whom), the time during which NO
kernel runs on the device and the time during which both streams run.

    python scripts/step_timeline.py [config] [out-prefix]        (default c2, gpurun_out/timeline)"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mmrseg_b200.losses import DiceCrossEntropyLoss
from mmrseg_b200.optim import FusedAdam

cfg = bench.resolve(sys.argv[1] if len(sys.argv) > 1 else "c2", 1)
prefix = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/timeline"
torch.manual_seed(6210)
model = bench.build_model(cfg, torch.device("cuda", 0)).train()
crit = DiceCrossEntropyLoss(0.5)
opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
x, y = bench.synthetic(cfg, cfg["batch"])
x, y = x.cuda(), y.cuda()


def step():
    for p in model.parameters():
        p.grad = None
    out = model(x)
    loss = sum(crit(o, y) for o in out) / len(out) if isinstance(out, list) else crit(out, y)
    loss.backward()
    opt.step()


for _ in range(6):
    step()
torch.cuda.synchronize()
NSTEP = 3
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(NSTEP):
        step()
    torch.cuda.synchronize()
trace = prefix + "_trace.json"
prof.export_chrome_trace(trace)
ev = [e for e in json.load(open(trace))["traceEvents"] if e.get("cat") == "kernel"]
os.remove(trace)
ev.sort(key=lambda e: e["ts"])
ks = [(e["name"].split("(")[0].replace("void ", "").replace("mmr::", ""), float(e["ts"]), float(e["dur"]),
       e["args"].get("stream")) for e in ev]
# steps are delimited by the weight-packing launch that opens every step
marks = [i for i, k in enumerate(ks) if "pack_weights_halo_batch" in k[0]]
lines = ["%d kernel records, %d steps found" % (len(ks), len(marks))]
with open(prefix + "_kernels.csv", "w") as f:
    f.write("name,stream,start_us,dur_us\n")
    lo = marks[1] if len(marks) > 1 else 0
    hi = marks[2] if len(marks) > 2 else len(ks)
    for n, ts, d, s in ks[lo:hi]:
        f.write("%s,%s,%.3f,%.3f\n" % (n.replace(",", ";"), s, ts - ks[lo][1], d))
for si in range(1, len(marks)):
    lo, hi = marks[si], (marks[si + 1] if si + 1 < len(marks) else len(ks))
    st = ks[lo:hi]
    t0 = st[0][1]
    t1 = max(ts + d for _, ts, d, _ in st)
    nxt = ks[hi][1] if hi < len(ks) else t1
    streams = collections.Counter(s for _, _, _, s in st)
    main = streams.most_common(1)[0][0]
    lines.append("step %d: %d kernels, first start -> last end %.1f us, next step starts %.1f us after this one's start; "
                 "streams %s" % (si, len(st), t1 - t0, nxt - t0, dict(streams)))
    # union coverage and two-stream overlap by a sweep
    pts = []
    for _, ts, d, _ in st:
        pts.append((ts, 1))
        pts.append((ts + d, -1))
    pts.sort()
    depth, last, idle, both = 0, t0, 0.0, 0.0
    for t, dlt in pts:
        if depth == 0:
            idle += t - last
        if depth >= 2:
            both += t - last
        depth += dlt
        last = t
    lines.append("  no kernel running: %.1f us; two or more kernels running: %.1f us" % (idle, both))
    for s in streams:
        sel = sorted((ts, d, n) for n, ts, d, ss in st if ss == s)
        busy = sum(d for _, d, _ in sel)
        gaps = []
        for (a_ts, a_d, a_n), (b_ts, b_d, b_n) in zip(sel, sel[1:]):
            gaps.append((b_ts - (a_ts + a_d), a_n, b_n))
        pos = [g for g in gaps if g[0] > 0]
        lines.append("  stream %s: %d kernels, busy %.1f us, %d gaps summing %.1f us (median %.2f us)" % (
            s, len(sel), busy, len(pos), sum(g[0] for g in pos),
            sorted(g[0] for g in pos)[len(pos) // 2] if pos else 0.0))
        if s == main:
            by = collections.defaultdict(lambda: [0.0, 0])
            for g, a_n, b_n in pos:
                k = (a_n[:34], b_n[:34])
                by[k][0] += g
                by[k][1] += 1
            for (a_n, b_n), (tot, cnt) in sorted(by.items(), key=lambda kv: -kv[1][0])[:14]:
                lines.append("      %-34s -> %-34s %3d gaps %7.1f us (%.2f each)" % (a_n, b_n, cnt, tot, tot / cnt))
    agg = collections.defaultdict(lambda: [0.0, 0])
    for n, ts, d, s in st:
        agg[n][0] += d
        agg[n][1] += 1
    if si == 1:
        lines.append("  kernel time inside the step (concurrent, graph replay):")
        for n, (tot, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
            lines.append("      %-52s %4d %9.1f us" % (n[:52], cnt, tot))
open(prefix + ".txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
