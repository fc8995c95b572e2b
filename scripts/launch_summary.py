"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches of ONE train step
(between two weight-packing launches) aggregated by kernel.
usage: python scripts/launch_summary.py file.csv [k]      k: which step of the file (default 0; the first one of a
process also holds the one-time fills of freshly allocated buffers)"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
seq = []
for x in csv.DictReader(lines):
    v = float(x["Metric Value"].replace(",", ""))
    v = v / 1e3 if x["Metric Unit"] == "ns" else (v * 1e3 if x["Metric Unit"] == "ms" else v)
    seq.append((x["Kernel Name"].split("(")[0].replace("void ", ""), v, x["Grid Size"]))
marks = [i for i, s in enumerate(seq) if "pack_weights_halo_batch" in s[0]]
print("launches in file: %d, step boundaries at %s" % (len(seq), marks))
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
k = min(k, max(len(marks) - 2, 0))
lo, hi = (marks[k], marks[k + 1]) if len(marks) > 1 else (0, len(seq))
step = seq[lo:hi]
agg = collections.defaultdict(lambda: [0.0, 0])
for n, v, g in step:
    agg[n][0] += v
    agg[n][1] += 1
tot = sum(v[0] for v in agg.values())
print("one step: %d launches, %.1f us of kernel time" % (len(step), tot))
for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-58s %9.1f us %5.1f %% %4d" % (n[:58], v, 100 * v / tot, c))
