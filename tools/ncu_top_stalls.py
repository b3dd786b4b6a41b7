"""Top stalled SASS instructions of an exported `ncu --page source --csv` file."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
print("total samples", tot)
for r in top:
    s = int(r[ix["# Samples"]] or 0)
    reasons = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print("%6d %5.1f%%  %-70s %s" % (s, 100.0 * s / tot, r[ix["Source"]].strip()[:70], " ".join("%s=%d" % (n, v) for v, n in reasons if v)))
