"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one forward step (between two consecutive
stem / ingest launches), grouped by kernel.  usage: python tools/launch_summary.py launches.csv [out_prefix]"""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    out = []
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)
        out.append((int(r["ID"]), r["Kernel Name"], r["Grid Size"], r["Block Size"], us))
    return out


def main():
    rows = load(sys.argv[1])
    starts = [i for i, r in enumerate(rows) if "stem_conv_u8" in r[1] or "ingest_u8" in r[1]]
    if len(starts) >= 2:
        step = rows[starts[-2]:starts[-1]]
    else:
        step = rows
    total = sum(r[4] for r in step)
    by = collections.OrderedDict()
    for _, name, g, b, us in step:
        k = name[:110]
        d = by.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += us
    print(f"# one step: {len(step)} launches, {total:.1f} us (cold-cache, serialised)")
    lines = ["kernel,launches,total_us,share"]
    for k, (n, us) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"\"{k}\",{n},{us:.1f},{us / total:.4f}")
    print("\n".join(lines[:45]))
    if len(sys.argv) > 2:
        with open(sys.argv[2] + "_by_kernel.csv", "w") as f:
            f.write("\n".join(lines) + "\n")
        with open(sys.argv[2] + ".csv", "w") as f:
            f.write("id,kernel,grid,block,us\n")
            for i, name, g, b, us in step:
                f.write(f"{i},\"{name[:160]}\",\"{g}\",\"{b}\",{us:.2f}\n")
    ours = sum(us for k, (n, us) in by.items() if any(ns in k for ns in ("el::", "pw::", "tc::")))
    print(f"# libedgeline_b200 kernels (el:: / pw:: / tc::) {ours:.1f} us = {ours / total:.3f} of the step")


if __name__ == "__main__":
    main()
