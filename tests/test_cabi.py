"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly what
include/edgeline_b200.h declares; argument validation happens before any CUDA work; the Python
host layer refuses CPU tensors (there is no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from edge_yolo_b200 import _lib, build

    build.build()
    return _lib.lib()


def _declared():
    text = open(os.path.join(ROOT, "include", "edgeline_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(el_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(lib):
    from edge_yolo_b200 import _lib

    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTED) == declared


def test_version_and_status_strings(lib):
    assert b"sm_100a" in lib.el_version()
    assert lib.el_status_string(0) == b"ok"
    assert lib.el_status_string(2) == b"unsupported shape"


def test_argument_validation_without_gpu(lib):
    z = (ctypes.c_int64 * 5)()
    assert lib.el_dwt_haar_fwd(None, z, None, z, 1, 1, 2, 2, 0, None) == 1
    assert lib.el_linattn_fwd(None, z, None, z, 1, 1, 4, 0, None) == 1
    n = ctypes.c_size_t()
    assert lib.el_nms_workspace_bytes(64, 80, 8400, 0, 30000, ctypes.byref(n)) == 0 and n.value > 64 * 8400 * 8
    assert lib.el_nms_workspace_bytes(64, 80, 8400, 1, 30000, ctypes.byref(n)) == 0 and n.value > 64 * 8400 * 80 * 8
    assert lib.el_nms_workspace_bytes(0, 80, 8400, 1, 30000, ctypes.byref(n)) == 1
    assert lib.el_nms_batched(None, 1, 1, 1, 0.5, 0.5, 0, 0, None, 300, 30000, 7680.0, None, 0, None, None, None, None) == 1
    assert lib.el_qfl_partials(64 * 8400 * 80) >= 148


def test_pwconv_refuses_in_place_output_when_n_is_tiled(lib):
    """el_pwconv_fwd: a destination aliasing a K source is a cross-CTA write-after-read race as soon as the output channels are split over
    several CTAs: refused with EL_ERR_ARG before any CUDA work; disjoint channel slices of one buffer stay legal (they
    fail later, on this GPU-less machine, with EL_ERR_CUDA from the tensor-map encode)."""
    from ctypes import c_int32, c_int64, c_void_p

    M, c = 64 * 20 * 20, 128
    assert -(-c // lib.el_pwconv_tile(c, 2 * 3 * c, M)) > 1  # the enhancer's fuse GEMM at 20x20, batch 64: two output-channel tiles
    b, U, w, fresh = 0x10000000, 0x20000000, 0x30000000, 0x40000000
    call = lambda srcs, pitches, cs, out, out_pitch, N=c: lib.el_pwconv_fwd(
        len(srcs), (c_void_p * len(srcs))(*srcs), (c_int64 * len(srcs))(*pitches), (c_int32 * len(srcs))(*cs), w, None, b, c, 1.0, 0, 0, out, out_pitch,
        None, 0, 0, M, N, 1, 2, None)
    assert call([b, U], [c, 2 * c], [c, 2 * c], b, c) == 1            # out aliases source 0: refused
    assert call([b, U], [c, 2 * c], [c, 2 * c], fresh, c) != 1        # out of place: passes validation
    # channel slices of ONE (M, 4c) buffer: source = channels [0, 2c), destination = channels [2c, 3c) -> legal; overlapping windows -> refused
    assert call([b], [4 * c], [2 * c], b + 2 * (2 * c), 4 * c) != 1
    assert call([b], [4 * c], [2 * c], b + 2 * c, 4 * c) == 1


def test_no_cpu_fallback():
    from edge_yolo_b200 import EdgelineError, ops

    x = torch.randn(1, 4, 8, 8)
    with pytest.raises(EdgelineError):
        ops.dwt_haar(x)
    with pytest.raises(EdgelineError):
        ops.linear_attention(torch.randn(1, 192, 4, 4), 1)
    with pytest.raises(EdgelineError):
        ops.nms_batched(torch.rand(1, 6, 10))
    from edge_yolo_b200.detection_loss import TaskAlignedAssigner

    with pytest.raises(EdgelineError):  # the default assigner is the kernel path: CPU tensors are an error, not a detour through torch ops
        TaskAlignedAssigner(topk=2, num_classes=3)(torch.rand(1, 4, 3), torch.rand(1, 4, 4), torch.rand(4, 2), torch.zeros(1, 1, 1),
                                                    torch.rand(1, 1, 4), torch.ones(1, 1, 1, dtype=torch.bool))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "edge_yolo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle: the product path must not depend on it"


def test_standalone_graph_matches_reference_parameter_counts():
    from edge_yolo_b200.model import EdgeLineYOLO

    # SURVEY.md appendix A (counted on the reference): n 2,678,699  s 9,617,483
    assert sum(p.numel() for p in EdgeLineYOLO("n", 80).parameters()) == 2678699
    assert sum(p.numel() for p in EdgeLineYOLO("s", 80).parameters()) == 9617483
    assert sum(p.numel() for p in EdgeLineYOLO("s", 10).parameters()) == 9590393


def test_nms_signature_matches_reference():
    import inspect

    from edge_yolo_b200.nms import non_max_suppression

    names = list(inspect.signature(non_max_suppression).parameters)
    assert names == ["prediction", "conf_thres", "iou_thres", "classes", "agnostic", "multi_label", "labels", "max_det", "nc",
                     "max_time_img", "max_nms", "max_wh", "in_place", "rotated"]
    with pytest.raises(AssertionError):
        non_max_suppression(torch.zeros(1, 6, 4), conf_thres=1.5)
    # end-to-end shaped input never touches a kernel (ops.py:224-228)
    out = non_max_suppression(torch.tensor([[[0, 0, 1, 1, 0.9, 1.0], [0, 0, 1, 1, 0.1, 2.0]]]), conf_thres=0.25)
    assert out[0].shape == (1, 6)


@pytest.mark.parametrize("N,src_c,M", [(64, [64], 1 << 20), (256, [128, 128], 25600), (80, [16, 48], 6400), (48, [32, 8, 72], 1 << 20)])
def test_pack_pw_weight_layout(N, src_c, M):
    """Host-side packing of a 1x1 conv weight into the resident UMMA B operand (ops.pack_pw_weight) against an independent reader of
    the documented layout: per output-channel tile and K chunk one [n_tile][box] K-major tile, 16-byte chunk j of row r stored at
    chunk j ^ ((r * row_bytes >> 7) & (row_bytes / 16 - 1)) (the TMA / UMMA 32-, 64- and 128-byte swizzles), tiles padded to 1 KiB."""
    from edge_yolo_b200 import ops

    K = sum(src_c)
    w = ((torch.arange(N * K) * 7) % 251 - 125).float().reshape(N, K)      # exactly representable in bf16
    from edge_yolo_b200 import _lib

    n_tile = _lib.lib().el_pwconv_tile(N, 2 * sum(b for _, _, b in ops._pw_chunks(src_c)), M)
    assert n_tile > 0 and n_tile % 16 == 0
    packed = ops.pack_pw_weight(w, src_c, torch.bfloat16, M).float().numpy()
    got = np.zeros((-(-N // n_tile) * n_tile, K), np.float32)
    pos = 0
    for t in range(-(-N // n_tile)):
        for k0, creal, bw in ops._pw_chunks(src_c):
            rb = 2 * bw                                                    # row bytes = swizzle span
            tile = packed[pos : pos + n_tile * bw].reshape(n_tile, bw // 8, 8)
            for r in range(n_tile):
                f = ((r * rb) >> 7) & (rb // 16 - 1)
                for j in range(bw // 8):
                    c0 = 8 * j
                    if c0 < creal:
                        got[t * n_tile + r, k0 + c0 : k0 + c0 + 8] = tile[r, j ^ f]
                    else:
                        assert not tile[r, j ^ f].any()                    # channel padding of a narrow last box is zero
            pos += -(-(n_tile * bw * 2) // 1024) * 512                     # elements: tile bytes rounded up to 1 KiB
    assert pos == packed.size
    np.testing.assert_array_equal(got[:N], w.numpy())
    assert not got[N:].any()                                               # output-channel padding rows are zero
