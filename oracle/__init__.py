"""CPU oracle for the face-detection hot path of sgasse/infercam_onnx.

THIS PACKAGE IS TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` leg may import it, and only as the
checker or as the timed CPU baseline — never as part of the product path. The
product (infercam_onnx_b200/) does not import it and fails loudly if its CUDA
library is missing.

Parity status: "parity unpinned" — see the headers of hotpath_oracle.c and
ultraface_ref.py and DESIGN.md §oracle.
"""
