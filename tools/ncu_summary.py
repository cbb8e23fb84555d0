"""Per-launch summary of an `ncu --set full` report (ncu -i rep --page raw --csv > raw.csv): python tools/ncu_summary.py raw.csv [out.csv]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"), ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wave_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct")]
have = [(h, n) for h, n in cols if h in ix]


def to_bytes(v, unit):
    v = float(v.replace(",", "")) if v not in ("", "n/a") else 0.0
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(unit, 1)


out = [",".join(n for _, n in have) + ",dram_MB,GBps_dram"]
for r in rows[2:]:
    vals = []
    rd = wr = us = 0.0
    for h, n in have:
        v = r[ix[h]]
        if n == "kernel":
            v = '"' + v[:90].replace('"', "'") + '"'
        elif n == "us":
            us = to_us(v, units[ix[h]])
            v = f"{us:.2f}"
        elif n == "dram_rd":
            rd = to_bytes(v, units[ix[h]])
            v = f"{rd / 1e6:.2f}"
        elif n == "dram_wr":
            wr = to_bytes(v, units[ix[h]])
            v = f"{wr / 1e6:.2f}"
        else:
            try:
                v = f"{float(v.replace(',', '')):.1f}"
            except ValueError:
                pass
        vals.append(v)
    vals += [f"{(rd + wr) / 1e6:.1f}", f"{(rd + wr) / us / 1e3:.0f}" if us else ""]
    out.append(",".join(vals))
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
