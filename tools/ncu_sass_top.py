"""Top stalled SASS instructions of the first kernel in an `ncu --page source --csv` dump, with the dominant stall reason.
    ncu -i x.ncu-rep --page source --csv > /tmp/x.csv ; python tools/ncu_sass_top.py /tmp/x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        body.append(r)
S = hdr.index("# Samples")
src = hdr.index("Source")
ie = hdr.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[S]) for r in body) or 1
print(f"{len(body)} SASS instructions, {tot} samples")
agg = {}
for r in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda i: -int(body[i][S]))[:n]
for i in sorted(order):
    r = body[i]
    top = max(stall_cols, key=lambda c: int(r[c] or 0))
    print(f"  #{i:5d} {100 * int(r[S]) / tot:5.1f}%  exec {int(r[ie]):7d}  {hdr[top][6:]:14s} {r[src].strip()[:90]}")
