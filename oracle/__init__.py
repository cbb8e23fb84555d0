"""TEST INFRASTRUCTURE ONLY.

CPU restatement (torch fp32 / numpy / plain C) of the EdgeLine-YOLO custom-operator
path of OneWalkman/EDGE-YOLO.  It exists to *check* the CUDA product path and to
serve as the reported CPU baseline; it is never the thing shipped or measured as the
product.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  `edge_yolo_b200/` must not.

Parity status: PINNED.  Every function in `hotpath.py` / `nms_ref.c` is checked
against outputs of the unmodified reference run in the dev container
(`oracle/gen_golden.py` -> `tests/golden/*.npz`, asserted by
`tests/test_oracle_golden.py`).  The reference's own test-suite holds no vectors for
this path, so reference-generated fixtures are the pin.
"""
