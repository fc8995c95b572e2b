"""profiles/conv_traffic.json from an ncu pass over the conv fprop + dgrad launches of one train step:
    MMR_NO_GRAPH=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \\
        -k regex:'conv_halo_kernel|conv_gemm_tc' -s 170 -c 85 --csv --log-file gpurun_out/conv_traffic.csv \\
        python scripts/profile_step.py 16 4
bench.py reports `dram_bytes_per_launch` as roofline.traffic (average over the 85 launches, like `achieved`)."""
import csv
import json
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_traffic.csv"
lines = [l for l in open(src) if l.startswith('"')]
per = {}
for x in csv.DictReader(lines):
    v = float(x["Metric Value"].replace(",", ""))
    unit = x["Metric Unit"]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}.get(unit, 1)
    per.setdefault(x["ID"], {})[x["Metric Name"]] = v * mult
n = len(per)
rd = sum(p["dram__bytes_read.sum"] for p in per.values())
wr = sum(p["dram__bytes_write.sum"] for p in per.values())
t = sum(p["gpu__time_duration.sum"] for p in per.values())
out = {"launches": n, "dram_bytes_per_launch": (rd + wr) / n, "dram_read_bytes_per_step": rd,
       "dram_write_bytes_per_step": wr, "kernel_seconds_under_ncu": t,
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the conv_halo_kernel + "
                 "conv_gemm_tc_kernel launches of one U-Net++ train step, batch 16 @ 512x512 (scripts/conv_traffic.py)"}
json.dump(out, open("profiles/conv_traffic.json", "w"), indent=1)
print(out)
