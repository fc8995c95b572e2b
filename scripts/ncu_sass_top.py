"""Top SASS instructions of an `ncu --page source --csv` dump by stall samples / shared-memory conflicts.
usage: ncu -i rep --page source --csv > x.csv; python scripts/ncu_sass_top.py x.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
col = {name: i for i, name in enumerate(H)}
body = [r for r in rows[hdr + 1:] if len(r) == len(H)]


def f(r, name):
    try:
        return float(r[col[name]])
    except ValueError:
        return 0.0


tot = sum(f(r, "# Samples") for r in body)
print("instructions %d, samples %d" % (len(body), tot))
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in body) for s in stalls}
print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print("--- by samples")
for i, r in sorted(enumerate(body), key=lambda ir: -f(ir[1], "# Samples"))[:n]:
    top = sorted(stalls, key=lambda s: -f(r, s))[:2]
    print("%5d %6.2f%% ex %9d  %-70s %s" % (i, 100 * f(r, "# Samples") / max(tot, 1), f(r, "Instructions Executed"),
                                          r[col["Source"]][:70], " ".join("%s=%d" % (s[6:], f(r, s)) for s in top)))
print("--- by excessive shared wavefronts")
for i, r in sorted(enumerate(body), key=lambda ir: -f(ir[1], "L1 Wavefronts Shared Excessive"))[:10]:
    print("%5d exc %9d of %9d  %s" % (i, f(r, "L1 Wavefronts Shared Excessive"), f(r, "L1 Wavefronts Shared"), r[col["Source"]][:80]))
