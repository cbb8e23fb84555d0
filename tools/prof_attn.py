"""Linear-attention core in isolation: parity against the CPU oracle and CUDA-event time at the BASELINE shapes (rotating inputs > L2).

    python tools/prof_attn.py                      # new dispatch (TMA kernel for N <= 512)
    EL_LINATTN_NO_TMA=1 python tools/prof_attn.py  # the round-1 tcgen05 kernel
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402
from oracle import hotpath as O  # noqa: E402

DEV = "cuda"
PEAK = 6544.0


def time_op(fn, sets, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(3_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in sets:
            fn(s)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3 / len(sets))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    out = {"kernel": "round-1 tcgen05" if os.environ.get("EL_LINATTN_NO_TMA") else "TMA chunk-parallel (N <= 512) / round-1 (N > 512)", "shapes": []}
    for name, (B, heads, hw), dtype in [("configs[1] n b64", (64, 2, (20, 20)), torch.bfloat16), ("configs[3] m b512", (512, 4, (20, 20)), torch.bfloat16),
                                        ("configs[2] s@1280 b32", (32, 4, (40, 40)), torch.bfloat16), ("s b64", (64, 4, (20, 20)), torch.bfloat16),
                                        ("configs[1] fp16", (64, 2, (20, 20)), torch.float16)]:
        C = heads * 64
        gen = torch.Generator().manual_seed(5)
        qkv = (torch.randn(B, 3 * C, *hw, generator=gen) * 1.5).to(dtype)
        x = qkv.to(DEV).contiguous(memory_format=torch.channels_last)
        y = ops.linear_attention(x, heads)
        nb = min(B, 4)
        ref = O.linear_attention_core(qkv[:nb].float(), heads)
        err = float((y[:nb].float().cpu() - ref).abs().max() / ref.abs().max())
        nbytes = 4 * B * C * hw[0] * hw[1] * 2
        R = max(2, min(48, (600 << 20) // nbytes))
        sets = [x] + [x.clone() for _ in range(R - 1)]
        t = time_op(lambda s: ops.linear_attention(s, heads), sets)
        flops = 4 * B * hw[0] * hw[1] * 64 * C
        out["shapes"].append({"name": name, "B": B, "heads": heads, "N": hw[0] * hw[1], "dtype": str(dtype), "rel_err_vs_oracle": err, "us": t * 1e6,
                              "MB": nbytes / 1e6, "GBs": nbytes / t / 1e9, "frac_of_hbm_peak": nbytes / t / 1e9 / PEAK, "TFLOPs": flops / t / 1e12})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
