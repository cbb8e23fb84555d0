"""Builds libedgeline_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

`python -m edge_yolo_b200.build` or `__graft_entry__.build()`.  The .so is git-ignored but ships
to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libedgeline_b200.so")
SOURCES = ["core.cu", "wavelet.cu", "linattn.cu", "linattn_tc.cu", "linattn_tma.cu", "linattn_bwd.cu", "decode.cu", "nms.cu", "loss.cu", "epilogue.cu", "dwconv.cu", "dwconv_tma.cu", "pwconv.cu", "dsconv.cu", "conv3x3_halo.cu", "conv3x3_mma.cu", "stem_tc.cu", "metrics.cu", "tal.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# no --use_fast_math: IEEE division / sqrt are part of the NMS bit-exactness contract
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _stale(out, deps):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "edgeline_b200.h"))

    def compile_one(src):
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src + ".o")
        if force or _stale(o, [s] + headers):
            cmd = [NVCC, *FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
