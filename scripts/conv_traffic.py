"""profiles/conv_traffic.json + a per-kernel DRAM-traffic table of ONE train step, from an ncu pass over every launch:
    MMR_NO_GRAPH=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \\
        -s 640 -c 760 --csv --log-file gpurun_out/step_traffic.csv python scripts/profile_step.py 16 4
    python scripts/conv_traffic.py gpurun_out/step_traffic.csv [table.txt]
The step is the stretch between two weight-packing launches (like scripts/launch_summary.py).  bench.py reports
`dram_bytes_per_launch` as roofline.traffic: the DRAM bytes of every tcgen05 conv kernel of the step (fprop, dgrad,
wgrad and the split-K partial reduce that belongs to a wgrad launch) divided by the number of conv plan launches
(fprop + dgrad + wgrad), i.e. per launch like roofline.achieved."""
import collections
import csv
import json
import re
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_traffic.csv"
lines = [l for l in open(src) if l.startswith('"')]
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
per = collections.OrderedDict()
for x in csv.DictReader(lines):
    v = float(x["Metric Value"].replace(",", ""))
    e = per.setdefault(int(x["ID"]), {"name": x["Kernel Name"].split("(")[0].replace("void ", "")})
    e[x["Metric Name"]] = v * MULT.get(x["Metric Unit"], 1)
seq = [per[k] for k in sorted(per)]
marks = [i for i, s in enumerate(seq) if "pack_weights_halo_batch" in s["name"]]
lo, hi = (marks[0], marks[1]) if len(marks) > 1 else (0, len(seq))
step = seq[lo:hi]
CONV = re.compile(r"conv_halo_kernel|conv_gemm_tc_kernel|conv_wgrad|wgrad_halo_reduce_kernel|wgrad_reduce_kernel")
PLAN = re.compile(r"conv_halo_kernel|conv_gemm_tc_kernel|conv_wgrad_(kx|thin|halo|tc)_kernel")   # one per plan launch
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for s in step:
    a = agg[s["name"]]
    a[0] += s.get("dram__bytes_read.sum", 0.0)
    a[1] += s.get("dram__bytes_write.sum", 0.0)
    a[2] += s.get("gpu__time_duration.sum", 0.0)
    a[3] += 1
conv = [s for s in step if CONV.search(s["name"])]
n_plan = sum(1 for s in step if PLAN.search(s["name"]))
rd = sum(s.get("dram__bytes_read.sum", 0.0) for s in conv)
wr = sum(s.get("dram__bytes_write.sum", 0.0) for s in conv)
t = sum(s.get("gpu__time_duration.sum", 0.0) for s in conv)
out = {"launches": n_plan, "dram_bytes_per_launch": (rd + wr) / max(n_plan, 1), "dram_read_bytes_per_step": rd,
       "dram_write_bytes_per_step": wr, "kernel_seconds_under_ncu": t,
       "step_dram_bytes_all_kernels": sum(a[0] + a[1] for a in agg.values()),
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every launch of one U-Net++ train step, "
                 "batch 16 @ 512x512; conv = conv_halo_kernel + conv_gemm_tc_kernel + conv_wgrad_*_kernel + split-K "
                 "partial reduces, per conv plan launch (scripts/conv_traffic.py)"}
json.dump(out, open("profiles/conv_traffic.json", "w"), indent=1)
text = ["one step: %d launches, DRAM %.2f GB read + %.2f GB written, %.2f ms of kernel time under ncu" % (
    len(step), sum(a[0] for a in agg.values()) / 1e9, sum(a[1] for a in agg.values()) / 1e9,
    sum(a[2] for a in agg.values()) * 1e3),
    "%-52s %5s %10s %10s %9s %9s" % ("kernel", "n", "read MB", "write MB", "us", "GB/s")]
for n, a in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][1])):
    text.append("%-52s %5d %10.1f %10.1f %9.1f %9.0f" % (n[:52], a[3], a[0] / 1e6, a[1] / 1e6, a[2] * 1e6,
                                                        (a[0] + a[1]) / a[2] / 1e9 if a[2] > 0 else 0.0))
text.append("conv kernels: %d plan launches, %.1f MB per launch" % (n_plan, out["dram_bytes_per_launch"] / 1e6))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(text) + "\n")
print("\n".join(text))
