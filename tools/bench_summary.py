"""Print the per-kernel table of a bench.py JSON line.  usage: python tools/bench_summary.py bench.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.0f} {d['unit']}  {d['ms_per_step']:.3f} ms/step   e2e {d['e2e']['value']:.0f} (serial {d['e2e'].get('unpipelined_value', 0):.0f})  launches/step {d.get('gpu_launches_per_step')}")
print("roofline", d.get("roofline"))
print("cpu", d.get("cpu_baseline"))
tot = 0
for k, v in d.get("kernels", {}).items():
    tot += v["us"]
    print(f"  {k:18s} sites {v['launch_sites']:3d}  {v['us']:8.1f} us  {v['bytes'] / 1e6:8.1f} MB  {v['gbs']:7.1f} GB/s")
    if "-v" in sys.argv:
        for s in v["sites"]:
            print("      ", s)
print(f"  el kernels total {tot:.1f} us")
