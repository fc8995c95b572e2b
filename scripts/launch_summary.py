"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches of ONE train step
(between two weight-packing launches) aggregated by kernel.  usage: python scripts/launch_summary.py file.csv"""
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
lo, hi = (marks[0], marks[1]) if len(marks) > 1 else (0, len(seq))
step = seq[lo:hi]
agg = collections.defaultdict(lambda: [0.0, 0])
for n, v, g in step:
    agg[n][0] += v
    agg[n][1] += 1
tot = sum(v[0] for v in agg.values())
print("one step: %d launches, %.1f us of kernel time" % (len(step), tot))
for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-58s %9.1f us %5.1f %% %4d" % (n[:58], v, 100 * v / tot, c))
