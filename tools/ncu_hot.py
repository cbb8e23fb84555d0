"""Top stall sites of an `ncu --page source --csv` dump: python tools/ncu_hot.py src.csv [kernel_index] [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
hdr_i = starts[k]
end = starts[k + 1] - 1 if k + 1 < len(starts) else len(rows)
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[hdr_i + 1:end] if len(r) >= len(hdr)]
num = lambda v: int(v) if v and v.isdigit() else 0
tot = sum(num(r[ix["# Samples"]]) for r in data)
print(f"kernel {k} of {len(starts)}: total samples", tot)
agg = {s: sum(num(r[ix[s]]) for r in data) for s in stalls}
print({k_: v for k_, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:N]
for r in top:
    s = {k_: num(r[ix[k_]]) for k_ in stalls if num(r[ix[k_]])}
    best = sorted(s.items(), key=lambda kv: -kv[1])[:3]
    print(f"{r[ix['Address']][-5:]} {num(r[ix['# Samples']]):6d} {r[ix['Source']][:70]:70s} {best}")
