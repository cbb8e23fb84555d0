"""Writes tests/golden/wtconv.npz from the UNMODIFIED reference WTConv2d (nn/modules/conv.py:463-598).  Dev container only:
python -m oracle.gen_golden_wtconv"""
import os

import numpy as np
import torch

from . import ref_loader

CASES = {  # name: (B, C, H, W, kernel, levels, stride)
    "l1": (2, 8, 16, 16, 5, 1, 1),
    "odd_l2": (1, 4, 13, 11, 5, 2, 1),
    "l3_s2_k3": (2, 8, 24, 40, 3, 3, 2),
}


def main():
    ref_loader.load()
    from ultralytics.nn.modules.conv import WTConv2d

    out = {}
    for i, (name, (B, C, H, W, k, lv, st)) in enumerate(CASES.items()):
        torch.manual_seed(100 + i)
        m = WTConv2d(C, C, kernel_size=k, stride=st, wt_levels=lv).eval()
        with torch.no_grad():
            for p in m.parameters():
                if p.requires_grad:
                    p.copy_(torch.randn_like(p) * 0.3)
            x = torch.randn(B, C, H, W)
            y = m(x)
        out[f"{name}_x"] = x.numpy()
        out[f"{name}_y"] = y.numpy()
        out[f"{name}_cfg"] = np.array([k, lv, st])
        for key, v in m.state_dict().items():
            out[f"{name}_sd_{key}"] = v.numpy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wtconv.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), sorted(k for k in out if k.startswith("l1_sd_")))


if __name__ == "__main__":
    main()
