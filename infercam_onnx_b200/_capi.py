"""ctypes binding of include/ultraface_b200.h (the stub a Python consumer of the C ABI writes;
the Rust equivalent is rust/ultraface-sys/src/lib.rs)."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

UF_OK = 0
STATUS = {0: "UF_OK", 1: "UF_ERR_INVALID_ARG", 2: "UF_ERR_IO", 3: "UF_ERR_ONNX", 4: "UF_ERR_UNSUPPORTED",
          5: "UF_ERR_CUDA", 6: "UF_ERR_NO_DEVICE", 7: "UF_ERR_CAPACITY"}
UF_NORM_REFERENCE, UF_NORM_127_128 = 0, 1
UF_FLAG_FORCE_GENERIC, UF_FLAG_NO_GRAPH, UF_FLAG_NO_FUSION, UF_FLAG_NO_TC, UF_FLAG_FUSE_DW_TC, UF_FLAG_PDL = 1, 2, 4, 8, 16, 32
UF_FLAG_TMA_SIMT_PW = 64
UF_FLAG_DENSE3_TC = 128
UF_FLAG_NO_PRESTEM = 256
UF_FLAG_JPEG_HOST_HUFFMAN = 512


class uf_glyph(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("w", C.c_uint32), ("h", C.c_uint32), ("offset", C.c_uint32)]


class uf_det(C.Structure):
    _fields_ = [("x0", C.c_float), ("y0", C.c_float), ("x1", C.c_float), ("y1", C.c_float), ("conf", C.c_float)]


class uf_config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("onnx_path", C.c_char_p), ("net_w", C.c_uint32), ("net_h", C.c_uint32),
                ("max_iou", C.c_float), ("min_confidence", C.c_float), ("device", C.c_int32),
                ("max_batch", C.c_uint32), ("norm_preset", C.c_uint32), ("chunk", C.c_uint32), ("slots", C.c_uint32),
                ("resize_round_intermediate", C.c_uint32), ("flags", C.c_uint32), ("lanes", C.c_uint32), ("host_chunk", C.c_uint32)]


class uf_info(C.Structure):
    _fields_ = [("net_w", C.c_uint32), ("net_h", C.c_uint32), ("num_priors", C.c_uint32), ("num_layers", C.c_uint32),
                ("num_tensors", C.c_uint32), ("max_batch", C.c_uint32), ("chunk", C.c_uint32), ("slots", C.c_uint32),
                ("weight_bytes", C.c_uint64), ("workspace_bytes", C.c_uint64),
                ("algorithmic_bytes_per_frame", C.c_uint64), ("macs_per_frame", C.c_uint64)]


class uf_kernel_stat(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("launches", C.c_uint64), ("device_ms", C.c_double),
                ("algorithmic_bytes", C.c_uint64), ("compulsory_bytes", C.c_uint64), ("flops", C.c_uint64)]


class uf_jpeg_info(C.Structure):
    _fields_ = [("w", C.c_uint32), ("h", C.c_uint32), ("ncomp", C.c_uint32), ("hs", C.c_uint32 * 3), ("vs", C.c_uint32 * 3),
                ("nblocks", C.c_uint32), ("nonzero", C.c_uint32), ("quant", (C.c_uint16 * 64) * 3)]


class uf_batcher_config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("model", uf_config), ("devices", C.POINTER(C.c_int32)), ("n_devices", C.c_uint32),
                ("max_batch", C.c_uint32), ("max_delay_us", C.c_uint32), ("capacity", C.c_uint32), ("workers", C.c_uint32),
                ("det_cap", C.c_uint32), ("max_frame_bytes", C.c_uint32), ("annotate_quality", C.c_uint32), ("annotate_scale_w", C.c_float),
                ("annotate_scale_h", C.c_float), ("annotate_max_bytes", C.c_uint32)]


class uf_result(C.Structure):
    _fields_ = [("stream", C.c_uint64), ("user_tag", C.c_uint64), ("device", C.c_int32), ("status", C.c_int32),
                ("n_dets", C.c_uint32), ("batch_size", C.c_uint32), ("latency_us", C.c_uint64), ("file_bytes", C.c_uint32),
                ("reserved_", C.c_uint32)]


class uf_batcher_stats(C.Structure):
    _fields_ = [("submitted", C.c_uint64), ("dropped", C.c_uint64), ("completed", C.c_uint64), ("failed", C.c_uint64),
                ("batches", C.c_uint64)]


_p = C.POINTER
_void_pp = _p(C.c_void_p)
uf_batch_fn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, _p(C.c_void_p), _p(C.c_uint32), _p(C.c_uint32), C.c_uint32,
                          _p(uf_det), C.c_uint32, _p(C.c_uint32))
# name -> (restype, argtypes); every symbol include/ultraface_b200.h declares
SIGNATURES = {
    "uf_model_load": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_int32, C.c_uint32, _void_pp]),
    "uf_model_load_ex": (C.c_int, [_p(uf_config), _void_pp]),
    "uf_model_free": (None, [C.c_void_p]),
    "uf_model_info": (C.c_int, [C.c_void_p, _p(uf_info)]),
    "uf_infer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _p(uf_det), C.c_uint32, _p(C.c_uint32)]),
    "uf_infer_batch": (C.c_int, [C.c_void_p, _p(C.c_void_p), _p(C.c_uint32), _p(C.c_uint32), C.c_uint32, _p(uf_det),
                                 C.c_uint32, _p(C.c_uint32)]),
    "uf_infer_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _p(uf_det),
                                        C.c_uint32, _p(C.c_uint32)]),
    "uf_raw_outputs": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, _p(C.c_float), _p(C.c_float)]),
    "uf_preproc_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "uf_preproc_u8_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "uf_debug_prestem_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "uf_preproc_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _p(C.c_float)]),
    "uf_postproc": (C.c_int, [C.c_void_p, _p(C.c_float), _p(C.c_float), C.c_uint32, _p(uf_det), C.c_uint32,
                              _p(C.c_uint32), _p(C.c_int32)]),
    "uf_tensor_count": (C.c_int, [C.c_void_p, _p(C.c_uint32)]),
    "uf_tensor_info": (C.c_int, [C.c_void_p, C.c_uint32, _p(C.c_char_p), _p(C.c_uint32), _p(C.c_uint32), _p(C.c_uint32)]),
    "uf_tensor_read": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, _p(C.c_float)]),
    "uf_host_alloc": (C.c_int, [C.c_size_t, _void_pp]),
    "uf_host_free": (None, [C.c_void_p]),
    "uf_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "uf_profile_reset": (C.c_int, [C.c_void_p]),
    "uf_profile_read": (C.c_int, [C.c_void_p, _p(uf_kernel_stat), C.c_uint32, _p(C.c_uint32)]),
    "uf_launch_count": (C.c_int, [C.c_void_p, _p(C.c_uint64)]),
    "uf_debug_fail_after": (C.c_int, [C.c_void_p, C.c_int32]),
    "uf_onnx_inspect": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t, _p(C.c_size_t)]),
    "uf_resize_taps": (C.c_int, [C.c_uint32, C.c_uint32, _p(C.c_int32), _p(C.c_int32), _p(C.c_float), C.c_uint32,
                                 _p(C.c_uint32)]),
    "uf_infer_batch_jpeg": (C.c_int, [C.c_void_p, _p(C.c_void_p), _p(C.c_size_t), C.c_uint32, _p(uf_det), C.c_uint32, _p(C.c_uint32)]),
    "uf_jpeg_coefficients_gpu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _p(C.c_int32)]),
    "uf_jpeg_decode_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _p(C.c_uint32), _p(C.c_uint32)]),
    "uf_jpeg_info_read": (C.c_int, [C.c_void_p, C.c_size_t, _p(uf_jpeg_info)]),
    "uf_jpeg_coefficients": (C.c_int, [C.c_void_p, C.c_size_t, _p(uf_jpeg_info), C.c_void_p, C.c_size_t]),
    "uf_annotate_encode_jpeg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_float, C.c_float,
                                          C.c_uint32, C.c_void_p, C.c_size_t, _p(C.c_size_t)]),
    "uf_annotate_reencode_jpeg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_uint32,
                                            C.c_void_p, C.c_size_t, _p(C.c_size_t)]),
    "uf_jpeg_write_coefficients": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _p(C.c_size_t)]),
    "uf_jpeg_quality_tables": (C.c_int, [C.c_uint32, C.c_void_p, C.c_void_p]),
    "uf_annotate_reencode_batch_jpeg": (C.c_int, [C.c_void_p, _p(C.c_void_p), _p(C.c_size_t), C.c_uint32, C.c_void_p, _p(C.c_uint32),
                                                  C.c_float, C.c_float, C.c_uint32, C.c_void_p, C.c_size_t, _p(C.c_size_t)]),
    "uf_worker_batch_jpeg": (C.c_int, [C.c_void_p, _p(C.c_void_p), _p(C.c_size_t), C.c_uint32, C.c_float, C.c_float, C.c_uint32, C.c_void_p,
                                       C.c_uint32, _p(C.c_uint32), C.c_void_p, C.c_size_t, _p(C.c_size_t)]),
    "uf_text_atlas_set": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "uf_confidence_text": (C.c_int, [C.c_float, C.c_char_p, C.c_size_t]),
    "uf_draw_boxes_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_void_p]),
    "uf_batcher_create": (C.c_int, [_p(uf_batcher_config), _void_pp]),
    "uf_batcher_create_ex": (C.c_int, [_p(uf_batcher_config), uf_batch_fn, C.c_void_p, _void_pp]),
    "uf_batcher_destroy": (None, [C.c_void_p]),
    "uf_batcher_acquire": (C.c_int, [C.c_void_p, C.c_uint64, C.c_size_t, _void_pp, _p(C.c_uint64)]),
    "uf_batcher_commit": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]),
    "uf_batcher_abort": (C.c_int, [C.c_void_p, C.c_uint64]),
    "uf_batcher_try_submit": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, _p(C.c_int32)]),
    "uf_batcher_commit_jpeg": (C.c_int, [C.c_void_p, C.c_uint64, C.c_size_t, C.c_uint64]),
    "uf_batcher_try_submit_jpeg": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_uint64, _p(C.c_int32)]),
    "uf_batcher_ingest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, _p(C.c_int32), _p(C.c_uint64)]),
    "uf_batcher_poll": (C.c_int, [C.c_void_p, _p(uf_result), _p(uf_det), C.c_uint32, C.c_uint32, _p(C.c_uint32)]),
    "uf_batcher_poll_frames": (C.c_int, [C.c_void_p, _p(uf_result), _p(uf_det), C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, _p(C.c_uint32)]),
    "uf_batcher_flush": (C.c_int, [C.c_void_p, C.c_uint32]),
    "uf_batcher_stats_read": (C.c_int, [C.c_void_p, _p(uf_batcher_stats)]),
    "uf_batcher_owner": (C.c_int, [C.c_void_p, C.c_uint64, _p(C.c_int32)]),
    "uf_batcher_model": (C.c_int, [C.c_void_p, C.c_uint32, _void_pp]),
    "uf_debug_batcher_drive": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _p(C.c_uint64), C.c_uint32, C.c_uint64,
                                         C.c_uint32, _p(C.c_double), _p(C.c_uint64)]),
    "uf_debug_batcher_drive_msgs": (C.c_int, [C.c_void_p, _p(C.c_char_p), _p(C.c_size_t), C.c_uint32, C.c_uint64, C.c_uint32,
                                              _p(C.c_double), _p(C.c_uint64)]),
    "uf_stream_hash": (C.c_int, [C.c_char_p, C.c_size_t, _p(C.c_uint64)]),
    "uf_protomsg_parse": (C.c_int, [C.c_char_p, C.c_size_t, _p(C.c_uint32), _void_pp, _p(C.c_size_t), _void_pp, _p(C.c_size_t)]),
    "uf_debug_siphash": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_char_p, C.c_size_t, _p(C.c_uint64)]),
    "uf_last_error": (C.c_char_p, []),
    "uf_version": (C.c_char_p, []),
    "uf_device_count": (C.c_int, [_p(C.c_int32)]),
}

_LIB = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree CUDA library; raises if it is missing and cannot be built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise OSError(f"{path} is missing: run `python -m infercam_onnx_b200.build` (nvcc, sm_100a)")
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
