"""One eager step under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` -> per-kernel launch count,
device time and DRAM bytes.  python tools/step_traffic.py step.csv out_prefix   (writes <prefix>_by_kernel.csv and <prefix>.json)"""
import collections
import csv
import json
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "us": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    else:
        b = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d["rd" if "read" in r["Metric Name"] else "wr"] = b


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([\w:]+)", name)
    return m.group(1) if m else name[:40]


by = collections.OrderedDict()
for d in launch.values():
    k = short(d["name"])
    a = by.setdefault(k, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    a["launches"] += 1
    a["us"] += d["us"]
    a["dram_bytes"] += d["rd"] + d["wr"]
tot = sum(a["us"] for a in by.values())
out = ["kernel,launches,total_us,share,dram_MB,dram_GBps"]
for k, a in sorted(by.items(), key=lambda kv: -kv[1]["us"]):
    out.append(f"{k},{a['launches']},{a['us']:.1f},{a['us'] / tot:.4f},{a['dram_bytes'] / 1e6:.1f},{a['dram_bytes'] / a['us'] / 1e3:.0f}")
text = "\n".join(out) + "\n"
print(f"# one step: {len(launch)} launches, {tot:.1f} us (cold-cache, serialised)")
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2] + "_by_kernel.csv", "w").write(text)
    json.dump({k: {"launches": a["launches"], "us": round(a["us"], 1), "dram_bytes": int(a["dram_bytes"])} for k, a in by.items()},
              open(sys.argv[2] + ".json", "w"), indent=1)
