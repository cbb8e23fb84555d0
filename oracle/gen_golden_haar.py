"""TEST INFRASTRUCTURE ONLY -- golden vectors for the fork's plain Haar modules from the unmodified reference.

    python -m oracle.gen_golden_haar      (dev container: needs /root/reference)

`HaarDWT2D.forward` (nn/modules/block.py:225-259) runs as it is.  `IHaarDWT2D` (block.py:2714-2750) cannot be constructed in the reference
(SURVEY Q5), so there is nothing to record for the synthesis: it is defined as the inverse of the analysis and pinned by the round trip
idwt(dwt(x)) == x on the reference's own analysis outputs stored here.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "haar.npz")


def main():
    ref_loader.load()
    import ultralytics.nn.modules.block as block

    g = torch.Generator().manual_seed(11)
    dwt = block.HaarDWT2D()
    out = {}
    for tag, shape in {"a": (2, 6, 10, 12), "b": (1, 3, 4, 8), "c": (2, 8, 12, 12)}.items():
        x = torch.randn(*shape, generator=g)
        bands = dwt(x)
        out[f"{tag}_x"] = x.numpy()
        out[f"{tag}_bands"] = torch.stack(bands, 0).numpy()
    try:
        block.IHaarDWT2D()
        out["ihaar_constructible"] = np.array(1)
    except TypeError:
        out["ihaar_constructible"] = np.array(0)  # Q5, recorded so that the tests notice if the reference ever changes
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes; IHaarDWT2D constructible in the reference:", int(out["ihaar_constructible"]))


if __name__ == "__main__":
    main()
