"""Summarise an .ncu-rep (read here, on the CPU box): per launch the metrics the roofline discussion uses.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--source N]   # --source: top-N source lines by stall samples of the first launch"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
        "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
        "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_drain_per_warp_active.pct",
        "smsp__warp_issue_stalled_imc_miss_per_warp_active.pct", "smsp__warp_issue_stalled_misc_per_warp_active.pct"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    idx = {k: hdr.index(k) for k in KEYS if k in hdr}
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[name_i][:90])
        for k, i in idx.items():
            if r[i] not in ("", "0"):
                print(f"   {k:86s} {r[i]}")
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass" if "--sass" in sys.argv else "cuda"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        hdr = next((r for r in rows if "Source" in r), None)
        if hdr is None:
            print("no source page")
            return
        si = hdr.index("Source")
        samp = next((hdr.index(c) for c in hdr if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"), None)
        inst = next((hdr.index(c) for c in hdr if c.startswith("Instructions Executed")), None)
        body = [r for r in rows[rows.index(hdr) + 1:] if len(r) > max(si, samp or 0)]
        def num(x):
            try:
                return float(x.replace(",", ""))
            except ValueError:
                return 0.0
        body.sort(key=lambda r: -num(r[samp]))
        tot = sum(num(r[samp]) for r in body) or 1
        print(f"-- top {n} source lines by stall samples (total {tot:.0f})")
        for r in body[:n]:
            print(f"   {100 * num(r[samp]) / tot:5.1f}%  inst {r[inst] if inst else '':>10s}  {r[si].strip()[:150]}")


if __name__ == "__main__":
    main()
