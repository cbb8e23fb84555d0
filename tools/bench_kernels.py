"""Prints the per-kernel table of a bench.py JSON line: python tools/bench_kernels.py gpurun_out/bench42.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.0f} img/s  {d['ms_per_step']:.3f} ms/step   e2e {d['e2e']['value']:.0f}   launches/step {d.get('gpu_launches_per_step')}")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["us"]):
    print(f"{k:18s} sites={v['launch_sites']:3d} us={v['us']:8.1f} gbs={v['gbs']:7.0f}")
    for s in sorted(v["sites"], key=lambda s: -s["us"] * s.get("count", 1))[:top]:
        print("      ", s)
for k in ("roofline", "cpu_baseline", "clocks"):
    if k in d:
        print(k, {a: b for a, b in d[k].items() if a != "note"})
