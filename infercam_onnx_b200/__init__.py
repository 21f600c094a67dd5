"""infercam_onnx_b200 — B200-native face-detection hot path of sgasse/infercam_onnx.

The product is `libultraface_b200.so` (hand-written sm_100a CUDA kernels behind the C ABI in
include/ultraface_b200.h). This package holds the build recipe, the ctypes binding and a
Python mirror of the reference's `infer_server::nn` interface (nn.py). There is no CPU
fallback: loading a model without a CUDA device raises.
"""
from .nn import Bbox, InferModel, UltrafaceError, UltrafaceModel, UltrafaceVariant  # noqa: F401

__all__ = ["Bbox", "InferModel", "UltrafaceError", "UltrafaceModel", "UltrafaceVariant"]
