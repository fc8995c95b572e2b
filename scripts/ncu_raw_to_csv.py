"""`ncu -i X.ncu-rep --page raw --csv` transposed into one `metric,unit,value` line per metric (what profiles/ keeps of a
full capture).  usage: python scripts/ncu_raw_to_csv.py gpurun_out/X.ncu-rep profiles/X.csv"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w") as f:
    for k, vals in enumerate(rows[2:]):
        f.write("# launch %d\nmetric,unit,value\n" % k)
        for h, u, v in zip(hdr, units, vals):
            f.write("%s,%s,%s\n" % (h, u, v.replace(",", "")))
