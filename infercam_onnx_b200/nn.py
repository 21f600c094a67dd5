"""Python mirror of the reference's `infer_server::nn` public surface over the C ABI.

Same names, argument meaning and error behaviour as
/root/reference/infer_server/src/nn.rs — `Bbox` (nn.rs:12), `InferModel` (nn.rs:24-26),
`UltrafaceVariant` + `width_height` (nn.rs:29-42), `UltrafaceModel::new(variant, max_iou,
min_confidence)` (nn.rs:55) and `run(&RgbImage) -> Result<Vec<(Bbox, f32)>>` (nn.rs:178-186) —
so the parity tests read like the reference's own integration test
(infer_server/tests/integration_tests.rs). Rust's `anyhow::Error` becomes `UltrafaceError`.
Everything numeric happens in libultraface_b200.so on the GPU; this file only marshals.
"""
from __future__ import annotations

import ctypes as C
import enum
import threading
import json
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _capi

Bbox = Tuple[float, float, float, float]  # [x_top_left, y_top_left, x_bottom_right, y_bottom_right], relative


class UltrafaceError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{_capi.STATUS.get(code, code)}: {message}")
        self.code = code


class UltrafaceVariant(enum.Enum):
    """nn.rs:29-42"""
    W640H480 = (640, 480)
    W320H240 = (320, 240)

    def width_height(self) -> Tuple[int, int]:
        return self.value

    def model_file_name(self) -> str:
        """nn.rs:144-147"""
        return {UltrafaceVariant.W640H480: "ultraface-RFB-640.onnx", UltrafaceVariant.W320H240: "ultraface-RFB-320.onnx"}[self]


class InferModel:
    """nn.rs:24-26"""

    def run(self, input: np.ndarray) -> List[Tuple[Bbox, float]]:  # noqa: A002 - the reference's parameter name
        raise NotImplementedError


def _check(rc: int) -> None:
    if rc != _capi.UF_OK:
        raise UltrafaceError(rc, _capi.load().uf_last_error().decode(errors="replace"))


def default_model_path(variant: UltrafaceVariant) -> str:
    """nn.rs:149-157: `dirs::cache_dir()/infercam_onnx/<model_name>` (no download here: no network)."""
    cache = os.environ.get("XDG_CACHE_HOME") or os.path.join(os.path.expanduser("~"), ".cache")
    return os.path.join(cache, "infercam_onnx", variant.model_file_name())


class UltrafaceModel(InferModel):
    """Loaded Ultraface model, ready for inference with post-processing thresholds (nn.rs:44-67)."""

    def __init__(self, handle: int, variant_wh: Tuple[int, int], max_iou: float, min_confidence: float):
        self._h = C.c_void_p(handle)
        self.width, self.height = variant_wh
        self.max_iou, self.min_confidence = max_iou, min_confidence
        info = _capi.uf_info()
        _check(_capi.load().uf_model_info(self._h, C.byref(info)))
        self.info = info
        self.num_priors = int(info.num_priors)
        self.max_batch = int(info.max_batch)
        self._tls = threading.local()
        self._ptr_args = {}

    @classmethod
    def new(cls, variant: UltrafaceVariant, max_iou: float, min_confidence: float, *, onnx_path: Optional[str] = None,
            device: int = 0, max_batch: int = 1, norm_preset: int = _capi.UF_NORM_REFERENCE, chunk: int = 0,
            slots: int = 0, flags: int = 0, resize_round_intermediate: bool = False, lanes: int = 0, host_chunk: int = 0,
            size: Optional[Tuple[int, int]] = None) -> "UltrafaceModel":
        """`UltrafaceModel::new` (nn.rs:55). The keyword arguments are extra knobs the reference
        hard-codes (cache path, nn.rs:149-157) or does not have (device, batch); `size` overrides the
        variant's (width, height) for graphs exported at another resolution."""
        lib = _capi.load()
        wh = size or variant.width_height()
        path = onnx_path or default_model_path(variant)
        cfg = _capi.uf_config()
        cfg.struct_size = C.sizeof(_capi.uf_config)
        cfg.onnx_path = os.fsencode(path)
        cfg.net_w, cfg.net_h = wh
        cfg.max_iou, cfg.min_confidence = max_iou, min_confidence
        cfg.device, cfg.max_batch, cfg.norm_preset = device, max_batch, norm_preset
        cfg.chunk, cfg.slots, cfg.flags = chunk, slots, flags
        cfg.resize_round_intermediate = int(resize_round_intermediate)
        cfg.lanes = lanes
        cfg.host_chunk = host_chunk
        h = C.c_void_p()
        _check(lib.uf_model_load_ex(C.byref(cfg), C.byref(h)))
        return cls(h.value, wh, max_iou, min_confidence)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            for buf in getattr(self._tls, "out", {}).values():  # this thread's pinned result arrays
                buf.free()
            self._tls = threading.local()
            _capi.load().uf_model_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # Drop
        try:
            self.close()
        except Exception:
            pass

    # ---- InferModel::run (nn.rs:178-186)
    def run(self, input: np.ndarray, cap: int = 256) -> List[Tuple[Bbox, float]]:  # noqa: A002
        img = _as_rgb(input)
        h, w = img.shape[:2]
        out = (_capi.uf_det * cap)()
        n = C.c_uint32()
        _check(_capi.load().uf_infer(self._h, img.ctypes.data_as(C.c_void_p), w, h, out, cap, C.byref(n)))
        if n.value > cap:
            return self.run(img, cap=int(n.value))
        return [((d.x0, d.y0, d.x1, d.y1), d.conf) for d in out[: n.value]]

    # ---- batched calls
    def run_batch(self, frames: Sequence[np.ndarray], cap: int = 256) -> List[np.ndarray]:
        """Host frames (any sizes) -> per-frame [n_i, 5] arrays (x0, y0, x1, y1, conf)."""
        imgs = [_as_rgb(f) for f in frames]
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        ws = (C.c_uint32 * n)(*[im.shape[1] for im in imgs])
        hs = (C.c_uint32 * n)(*[im.shape[0] for im in imgs])
        return self._run_batch_raw(lambda out, cnt: _capi.load().uf_infer_batch(self._h, ptrs, ws, hs, n, out, cap, cnt), n, cap)

    def run_batch_ptr(self, host_ptr: int, w: int, h: int, n: int, cap: int = 256) -> List[np.ndarray]:
        """n frames of w x h contiguous in (pinned) HOST memory at host_ptr."""
        key = (host_ptr, w, h, n)
        args = self._ptr_args.get(key)
        if args is None:  # the argument arrays of a repeated call (a ring of pinned frames) are built once
            fb = w * h * 3
            args = ((C.c_void_p * n)(*[host_ptr + i * fb for i in range(n)]), (C.c_uint32 * n)(*([w] * n)), (C.c_uint32 * n)(*([h] * n)))
            if len(self._ptr_args) < 64:
                self._ptr_args[key] = args
        ptrs, ws, hs = args
        return self._run_batch_raw(lambda out, cnt: _capi.load().uf_infer_batch(self._h, ptrs, ws, hs, n, out, cap, cnt), n, cap)

    def run_batch_device(self, device_ptr: int, w: int, h: int, n: int, cap: int = 256) -> List[np.ndarray]:
        """n frames of w x h contiguous in DEVICE memory at device_ptr."""
        return self._run_batch_raw(
            lambda out, cnt: _capi.load().uf_infer_batch_device(self._h, C.c_void_p(device_ptr), w, h, n, out, cap, cnt), n, cap)

    def run_batch_jpeg(self, jpegs: Sequence[bytes], cap: int = 256) -> List[np.ndarray]:
        """N2: frames as baseline JPEG files (bytes), decoded on the GPU (Huffman decoding included; host fallback for frames
        with restart intervals or damaged data)."""
        n = len(jpegs)
        ptrs, lens, keep = _jpeg_args(jpegs)
        out = self._run_batch_raw(lambda out, cnt: _capi.load().uf_infer_batch_jpeg(self._h, ptrs, lens, n, out, cap, cnt), n, cap)
        del keep
        return out

    def jpeg_decode_rgb(self, jpeg: bytes) -> np.ndarray:
        """Parity hook: the RGB8 pixels the GPU decode kernels produce for one JPEG file."""
        info = jpeg_info(jpeg)
        out = np.empty((info["h"], info["w"], 3), np.uint8)
        w, h = C.c_uint32(), C.c_uint32()
        buf = C.create_string_buffer(bytes(jpeg), len(jpeg))
        _check(_capi.load().uf_jpeg_decode_rgb(self._h, buf, len(jpeg), out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(w), C.byref(h)))
        return out

    def jpeg_coefficients_gpu(self, jpeg: bytes):
        """Parity hook: (coefficients [nblocks, 64] int16 as the DEVICE Huffman decoder produces them, launches): launches = 0
        if the frame was handed to the host decoder (restart intervals, damaged data), else the synchronisation launches."""
        info = jpeg_info(jpeg)
        out = np.zeros((info["nblocks"], 64), np.int16)
        on_dev = C.c_int32()
        buf = C.create_string_buffer(bytes(jpeg), len(jpeg))
        _check(_capi.load().uf_jpeg_coefficients_gpu(self._h, buf, len(jpeg), out.ctypes.data_as(C.c_void_p), out.shape[0], C.byref(on_dev)))
        return out, max(on_dev.value - 1, 0) if on_dev.value else 0

    # ---- N3 (rectangles + JPEG encode; inferer.rs:38-39, 58-92)
    @staticmethod
    def _dets_array(dets):
        d = np.ascontiguousarray(np.asarray(dets, np.float32).reshape(-1, 5))
        return d, d.ctypes.data_as(C.c_void_p), len(d)

    def annotate_reencode_batch_jpeg(self, jpegs: Sequence[bytes], dets_per_frame, scale_w: float, scale_h: float, quality: int = 95,
                                     out_stride: int = 0) -> List[bytes]:
        """Batch form of annotate_encode_jpeg for JPEG input: decode, overlay and encode on the GPU (entropy coding included)."""
        n = len(jpegs)
        ptrs, lens, keep = _jpeg_args(jpegs)
        flat = [np.asarray(d, np.float32).reshape(-1, 5) for d in dets_per_frame]
        counts = (C.c_uint32 * max(n, 1))(*[len(d) for d in flat])
        alld = np.ascontiguousarray(np.concatenate(flat) if flat else np.zeros((0, 5), np.float32))
        if not out_stride:
            out_stride = max(len(j) for j in jpegs) * 8 + (1 << 16) if n else 1024
        out = np.empty(max(n, 1) * out_stride, np.uint8)
        out_len = (C.c_size_t * max(n, 1))()
        _check(_capi.load().uf_annotate_reencode_batch_jpeg(self._h, ptrs, lens, n, alld.ctypes.data_as(C.c_void_p), counts, scale_w, scale_h,
                                                            quality, out.ctypes.data_as(C.c_void_p), out_stride, out_len))
        return [out[i * out_stride:i * out_stride + out_len[i]].tobytes() for i in range(n)]

    def worker_batch_jpeg(self, jpegs: Sequence[bytes], scale_w: float, scale_h: float, quality: int = 95, cap: int = 64, out_stride: int = 0,
                          keep_files: bool = True):
        """The reference's worker loop body for a batch (uf_worker_batch_jpeg): JPEG in -> (detections per frame, counts,
        annotated JPEG files)."""
        n = len(jpegs)
        key = ("worker", n, tuple(map(len, jpegs[:4])))
        cached = getattr(self, "_worker_args", None)
        if cached is None or cached[0] != key or cached[1] is not jpegs:
            ptrs, lens, bufs = _jpeg_args(jpegs)
            self._worker_args = cached = (key, jpegs, bufs, ptrs, lens)
        _, _, bufs, ptrs, lens = cached
        if not out_stride:
            out_stride = max(len(j) for j in jpegs) * 8 + (1 << 16) if n else 1024
        out = np.empty(max(n, 1) * out_stride, np.uint8)
        out_len = (C.c_size_t * max(n, 1))()
        dets = np.zeros((max(n, 1), cap, 5), np.float32)
        cnt = (C.c_uint32 * max(n, 1))()
        _check(_capi.load().uf_worker_batch_jpeg(self._h, ptrs, lens, n, scale_w, scale_h, quality, dets.ctypes.data_as(C.c_void_p), cap, cnt,
                                                 out.ctypes.data_as(C.c_void_p), out_stride, out_len))
        counts = [int(cnt[i]) for i in range(n)]
        files = [out[i * out_stride:i * out_stride + out_len[i]].tobytes() for i in range(n)] if keep_files else [int(out_len[i]) for i in range(n)]
        return [dets[i, :min(counts[i], cap)] for i in range(n)], counts, files

    def text_atlas_set(self, charset: str, max_len: int, glyphs, coverage) -> None:
        """Glyph atlas for the confidence text (uf_text_atlas_set): glyphs[pos][k] = (x0, y0, w, h, offset) of charset[k] as the
        pos-th character, coverage = flat f32 array. charset "" removes the atlas (rectangles only)."""
        if not charset:
            _check(_capi.load().uf_text_atlas_set(self._h, None, 0, 0, None, None, 0))
            return
        g = np.asarray(glyphs, np.int64).reshape(max_len * len(charset), 5)
        arr = (_capi.uf_glyph * len(g))(*[_capi.uf_glyph(int(a), int(b), int(c), int(d), int(e)) for a, b, c, d, e in g])
        cov = np.ascontiguousarray(coverage, np.float32).ravel()
        _check(_capi.load().uf_text_atlas_set(self._h, charset.encode("ascii"), len(charset), max_len, arr, cov.ctypes.data_as(C.c_void_p), cov.size))

    def draw_boxes(self, rgb: np.ndarray, dets, scale_w: float, scale_h: float) -> np.ndarray:
        img = _as_rgb(rgb)
        out = np.empty_like(img)
        d, dp, n = self._dets_array(dets)
        _check(_capi.load().uf_draw_boxes_rgb(self._h, img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], dp, n, scale_w, scale_h,
                                              out.ctypes.data_as(C.c_void_p)))
        return out

    def annotate_encode_jpeg(self, frame, dets, scale_w: float, scale_h: float, quality: int = 95) -> bytes:
        """frame: HxWx3 u8 RGB array, or the bytes of a baseline JPEG (decoded on the GPU). Returns a JPEG file."""
        d, dp, n = self._dets_array(dets)
        need = C.c_size_t()
        if isinstance(frame, (bytes, bytearray)):
            src = C.create_string_buffer(bytes(frame), len(frame))
            info = jpeg_info(bytes(frame))
            cap = info["w"] * info["h"] * 3 + 4096
            out = C.create_string_buffer(cap)
            _check(_capi.load().uf_annotate_reencode_jpeg(self._h, src, len(frame), dp, n, scale_w, scale_h, quality, out, cap, C.byref(need)))
        else:
            img = _as_rgb(frame)
            cap = img.size + 4096
            out = C.create_string_buffer(cap)
            _check(_capi.load().uf_annotate_encode_jpeg(self._h, img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], dp, n, scale_w,
                                                        scale_h, quality, out, cap, C.byref(need)))
        return out.raw[: need.value]

    def _run_batch_raw(self, call, n: int, cap: int) -> List[np.ndarray]:
        # result arrays live in pinned host memory, one set per calling thread and (n, cap), reused from call to call:
        # the library writes detections beyond the first 128 of a frame with an asynchronous strided copy, which is only
        # asynchronous (and fast) into pinned memory, and a 40 MB np.zeros per call would dominate an NMS-heavy batch
        tl = self._tls
        cache = getattr(tl, "out", None)
        if cache is None:
            cache = tl.out = {}
        key = (max(n, 1), cap)
        if key not in cache:
            if len(cache) >= 4:
                cache.pop(next(iter(cache))).free()
            cache[key] = _PinnedResults(*key)
        buf = cache[key]
        _check(call(buf.dets_p, buf.counts_p))
        counts = buf.counts[:n].tolist()
        # one copy out of the pinned array (only as deep as the fullest frame), per-frame results are views of it
        deepest = min(max(counts, default=0), cap)
        out = buf.dets[:n, :deepest].copy()
        return [out[i, : min(c, cap)] for i, c in enumerate(counts)], counts

    # ---- parity hooks
    def raw_outputs(self, first: int, n: int) -> Tuple[np.ndarray, np.ndarray]:
        s = np.empty((n, self.num_priors, 2), np.float32)
        b = np.empty((n, self.num_priors, 4), np.float32)
        _check(_capi.load().uf_raw_outputs(self._h, first, n, s.ctypes.data_as(C.POINTER(C.c_float)),
                                           b.ctypes.data_as(C.POINTER(C.c_float))))
        return s, b

    def preproc_u8(self, input: np.ndarray) -> np.ndarray:  # noqa: A002
        img = _as_rgb(input)
        out = np.empty((self.height, self.width, 3), np.uint8)
        _check(_capi.load().uf_preproc_u8(self._h, img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0],
                                          out.ctypes.data_as(C.c_void_p)))
        return out

    def preproc_u8_batch(self, frames: np.ndarray) -> np.ndarray:
        """nn.rs:74-80 for n same-size frames [n,h,w,3] in ONE launch (n <= the stage size) -> u8 [n,H,W,3]."""
        frames = np.ascontiguousarray(frames, np.uint8)
        if frames.ndim != 4 or frames.shape[3] != 3:
            raise ValueError("frames must be [n, h, w, 3] uint8")
        n, h, w = frames.shape[:3]
        out = np.empty((n, self.height, self.width, 3), np.uint8)
        _check(_capi.load().uf_preproc_u8_batch(self._h, frames.ctypes.data_as(C.c_void_p), w, h, n,
                                                out.ctypes.data_as(C.c_void_p)))
        return out

    def prestem_u8(self, frames: np.ndarray) -> np.ndarray:
        """The resized u8 pixels as seen by the fused resize+normalise+stem kernel ([n, 2H, 2W, 3] -> [n, H, W, 3])."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w = frames.shape[:3]
        out = np.empty((n, self.height, self.width, 3), np.uint8)
        _check(_capi.load().uf_debug_prestem_u8(self._h, frames.ctypes.data_as(C.c_void_p), w, h, n, out.ctypes.data_as(C.c_void_p)))
        return out

    def preproc(self, input: np.ndarray) -> np.ndarray:  # noqa: A002
        """nn.rs:70-94 -> f32 [1,3,H,W]"""
        img = _as_rgb(input)
        out = np.empty((1, 3, self.height, self.width), np.float32)
        _check(_capi.load().uf_preproc_f32(self._h, img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0],
                                           out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def postproc(self, scores: np.ndarray, boxes: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """nn.rs:109-140 on caller-supplied raw tensors -> ([n,5] detections, [n] prior indices)."""
        scores = np.ascontiguousarray(scores, np.float32).reshape(-1, 2)
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4)
        K = scores.shape[0]
        out = np.zeros((max(K, 1), 5), np.float32)
        idx = np.zeros(max(K, 1), np.int32)
        n = C.c_uint32()
        _check(_capi.load().uf_postproc(self._h, scores.ctypes.data_as(C.POINTER(C.c_float)),
                                        boxes.ctypes.data_as(C.POINTER(C.c_float)), K,
                                        out.ctypes.data_as(C.POINTER(_capi.uf_det)), K, C.byref(n),
                                        idx.ctypes.data_as(C.POINTER(C.c_int32))))
        return out[: n.value].copy(), idx[: n.value].copy()

    def tensors(self) -> dict:
        """ONNX value name -> (index, C, H, W) of every materialised activation tensor."""
        lib = _capi.load()
        n = C.c_uint32()
        _check(lib.uf_tensor_count(self._h, C.byref(n)))
        out = {}
        for i in range(n.value):
            name = C.c_char_p()
            c, h, w = C.c_uint32(), C.c_uint32(), C.c_uint32()
            _check(lib.uf_tensor_info(self._h, i, C.byref(name), C.byref(c), C.byref(h), C.byref(w)))
            if name.value:
                out[name.value.decode()] = (i, c.value, h.value, w.value)
        return out

    def tensor_read(self, index: int, frame: int, chw: Tuple[int, int, int]) -> np.ndarray:
        out = np.empty(chw, np.float32)
        _check(_capi.load().uf_tensor_read(self._h, index, frame, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    # ---- measurement
    def profile_enable(self, on: bool) -> None:
        _check(_capi.load().uf_profile_enable(self._h, int(on)))

    def profile_reset(self) -> None:
        _check(_capi.load().uf_profile_reset(self._h))

    def profile_read(self) -> List[dict]:
        arr = (_capi.uf_kernel_stat * 128)()
        n = C.c_uint32()
        _check(_capi.load().uf_profile_read(self._h, arr, 128, C.byref(n)))
        return [dict(name=a.name.decode(), launches=int(a.launches), device_ms=float(a.device_ms),
                     algorithmic_bytes=int(a.algorithmic_bytes), compulsory_bytes=int(a.compulsory_bytes),
                     flops=int(a.flops)) for a in arr[: min(n.value, 128)]]

    def profile_by_family(self) -> List[dict]:
        """profile_read() aggregated over the shapes of each kernel family."""
        fam = {}
        for s in self.profile_read():
            k = s["name"].split("[")[0]
            f = fam.setdefault(k, dict(name=k, launches=0, device_ms=0.0, algorithmic_bytes=0, compulsory_bytes=0, flops=0))
            for key in ("launches", "device_ms", "algorithmic_bytes", "compulsory_bytes", "flops"):
                f[key] += s[key]
        return list(fam.values())

    def debug_fail_after(self, stages: int) -> None:
        """Fault injection: following batched calls fail after `stages` pipeline stages (negative: off)."""
        _check(_capi.load().uf_debug_fail_after(self._h, stages))

    def launch_count(self) -> int:
        n = C.c_uint64()
        _check(_capi.load().uf_launch_count(self._h, C.byref(n)))
        return int(n.value)


def _jpeg_args(jpegs: Sequence[bytes]):
    """(array of pointers, array of lengths, what must stay alive during the call) for a list of JPEG files — pointers INTO the
    bytes objects, no copy of the files."""
    n = len(jpegs)
    keep = [j if isinstance(j, bytes) else bytes(j) for j in jpegs]
    cptrs = (C.c_char_p * max(n, 1))(*keep)
    lens = (C.c_size_t * max(n, 1))(*[len(j) for j in keep])
    return C.cast(cptrs, C.POINTER(C.c_void_p)), lens, (keep, cptrs)


def _as_rgb(a: np.ndarray) -> np.ndarray:
    """`&RgbImage` (nn.rs:25): contiguous HWC u8, any W x H."""
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise UltrafaceError(1, f"expected an HxWx3 uint8 RGB image, got {a.dtype} {a.shape}")
    return np.ascontiguousarray(a)


def onnx_inspect(path: str, width: int, height: int) -> dict:
    """Host-only: parse + lower an ONNX file and describe the launch plan (no GPU needed)."""
    lib = _capi.load()
    need = C.c_size_t()
    _check(lib.uf_onnx_inspect(os.fsencode(path), width, height, None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    _check(lib.uf_onnx_inspect(os.fsencode(path), width, height, buf, need.value, C.byref(need)))
    return json.loads(buf.value.decode())


def confidence_text(confidence: float) -> str:
    """The text the overlay prints for a confidence (Rust: format!("{:.2}%", confidence * 100.0) on f32). Host only."""
    buf = C.create_string_buffer(32)
    _check(_capi.load().uf_confidence_text(float(np.float32(confidence)), buf, 32))
    return buf.value.decode("ascii")


def jpeg_info(jpeg: bytes) -> dict:
    """Host-only: size / components / sampling factors from the JPEG headers."""
    info = _capi.uf_jpeg_info()
    buf = C.create_string_buffer(bytes(jpeg), len(jpeg))
    _check(_capi.load().uf_jpeg_info_read(buf, len(jpeg), C.byref(info)))
    nc = int(info.ncomp)
    return dict(w=int(info.w), h=int(info.h), ncomp=nc, hs=list(info.hs)[:nc], vs=list(info.vs)[:nc], nblocks=int(info.nblocks))


def jpeg_coefficients(jpeg: bytes):
    """Host-only: Huffman decoding alone -> (info dict incl. quant [ncomp,64] and the nonzero count, coefs [nblocks,64] int16;
    blocks in decode (MCU-interleaved) order, natural order inside a block)."""
    d = jpeg_info(jpeg)
    info = _capi.uf_jpeg_info()
    coefs = np.zeros((d["nblocks"], 64), np.int16)
    buf = C.create_string_buffer(bytes(jpeg), len(jpeg))
    _check(_capi.load().uf_jpeg_coefficients(buf, len(jpeg), C.byref(info), coefs.ctypes.data_as(C.c_void_p), d["nblocks"]))
    d["nonzero"] = int(info.nonzero)
    d["quant"] = np.ctypeslib.as_array(info.quant).reshape(3, 64)[: d["ncomp"]].copy()
    return d, coefs


def jpeg_write_coefficients(w: int, h: int, quality: int, coefs: np.ndarray) -> bytes:
    """Host-only: Huffman coding + file writing of quantised 4:2:0 blocks given per padded component plane in raster order."""
    coefs = np.ascontiguousarray(coefs, np.int16).reshape(-1, 64)
    cap = w * h * 3 + 65536
    out = C.create_string_buffer(cap)
    need = C.c_size_t()
    _check(_capi.load().uf_jpeg_write_coefficients(w, h, quality, coefs.ctypes.data_as(C.c_void_p), len(coefs), out, cap, C.byref(need)))
    return out.raw[: need.value]


def jpeg_quality_tables(quality: int):
    lum, chr_ = np.zeros(64, np.uint16), np.zeros(64, np.uint16)
    _check(_capi.load().uf_jpeg_quality_tables(quality, lum.ctypes.data_as(C.c_void_p), chr_.ctypes.data_as(C.c_void_p)))
    return lum, chr_


def resize_taps(src_len: int, dst_len: int):
    """Host-only: (left, ntaps, w[dst, max_taps]) of one resize axis as the GPU kernel will use them."""
    lib = _capi.load()
    mt = C.c_uint32()
    _check(lib.uf_resize_taps(src_len, dst_len, None, None, None, 0, C.byref(mt)))
    left = np.zeros(dst_len, np.int32)
    nt = np.zeros(dst_len, np.int32)
    w = np.zeros((dst_len, mt.value), np.float32)
    _check(lib.uf_resize_taps(src_len, dst_len, left.ctypes.data_as(C.POINTER(C.c_int32)),
                              nt.ctypes.data_as(C.POINTER(C.c_int32)), w.ctypes.data_as(C.POINTER(C.c_float)),
                              mt.value, C.byref(mt)))
    return left, nt, w


def device_count() -> int:
    n = C.c_int32()
    _check(_capi.load().uf_device_count(C.byref(n)))
    return int(n.value)


class _PinnedResults:
    """[n, cap, 5] f32 detections + [n] u32 counts in pinned host memory."""

    def __init__(self, n: int, cap: int):
        self.nbytes = n * cap * 20 + n * 4
        p = C.c_void_p()
        _check(_capi.load().uf_host_alloc(self.nbytes, C.byref(p)))
        self.ptr = p.value
        self.dets = np.ctypeslib.as_array((C.c_float * (n * cap * 5)).from_address(self.ptr)).reshape(n, cap, 5)
        self.counts = np.ctypeslib.as_array((C.c_uint32 * n).from_address(self.ptr + n * cap * 20))
        self.counts_p = C.cast(self.ptr + n * cap * 20, C.POINTER(C.c_uint32))
        self.dets_p = C.cast(self.ptr, C.POINTER(_capi.uf_det))

    def free(self) -> None:
        if self.ptr:
            self.dets = self.counts = None
            _capi.load().uf_host_free(C.c_void_p(self.ptr))
            self.ptr = None


class PinnedFrames:
    """n x h x w x 3 u8 frames in pinned host memory (uf_host_alloc), exposed as a numpy array."""

    def __init__(self, n: int, h: int, w: int):
        self.shape = (n, h, w, 3)
        self.nbytes = n * h * w * 3
        p = C.c_void_p()
        _check(_capi.load().uf_host_alloc(self.nbytes, C.byref(p)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr)).reshape(self.shape)

    def free(self) -> None:
        if self.ptr:
            self.array = None
            _capi.load().uf_host_free(C.c_void_p(self.ptr))
            self.ptr = None
