"""Stream batcher + stream -> GPU router — SURVEY.md §8f rows N1 / N4 — as a thin ctypes mirror of the C ABI
(`uf_batcher_*`, `uf_stream_hash`, `uf_protomsg_parse` in include/ultraface_b200.h; implementation csrc/batcher.cc).

The reference infers one frame at a time in one task (`Inferer::run`, /root/reference/infer_server/src/inferer.rs:29-50)
behind a bounded LOSSY queue (`INFER_IMAGES_CHANNEL`, capacity 10, lib.rs:32-37; `try_send_ref` drops the frame when it is
full, router.rs:64-72) and keys streams by `hashed(&id)` (lib.rs:39-46, router.rs:58). The C batcher keeps those
semantics and adds what a GPU needs: deadline-bounded batches, several batches in flight, one model handle per GPU with
`device = stream % n_devices`, per-stream ordered results. Everything below only marshals; the Rust server binds the
same symbols (rust/ultraface-sys).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from .nn import UltrafaceError, UltrafaceVariant, _as_rgb, _check, default_model_path


def stream_hash(name: str) -> int:
    """`hashed(&id)` of lib.rs:39-46 (Rust DefaultHasher = SipHash-1-3, zero keys, over the name bytes + 0xff)."""
    raw = name.encode()
    out = C.c_uint64()
    _check(_capi.load().uf_stream_hash(raw, len(raw), C.byref(out)))
    return int(out.value)


def protomsg_parse(msg: bytes) -> Tuple[str, str, bytes]:
    """One length-delimited data-socket frame (bincode `ProtoMsg`, common/src/protocol.rs:7-19) ->
    ("ConnectReq", id, b"") or ("FrameMsg", id, data)."""
    kind, idl, dl = C.c_uint32(), C.c_size_t(), C.c_size_t()
    idp, dp = C.c_void_p(), C.c_void_p()
    buf = C.create_string_buffer(bytes(msg), len(msg))
    _check(_capi.load().uf_protomsg_parse(buf, len(msg), C.byref(kind), C.byref(idp), C.byref(idl), C.byref(dp), C.byref(dl)))
    sid = C.string_at(idp.value, idl.value).decode() if idl.value else ""
    data = C.string_at(dp.value, dl.value) if dl.value else b""
    return ("ConnectReq", "FrameMsg")[kind.value], sid, data


class StreamBatcher:
    """`uf_batcher`: try_submit never blocks (False = dropped, like a full INFER_IMAGES_CHANNEL); results come back through
    poll() or, when `on_result` callbacks are used, from a dispatcher thread, per stream in submission order."""

    def __init__(self, variant: UltrafaceVariant = UltrafaceVariant.W320H240, max_iou: float = 0.5, min_confidence: float = 0.5, *,
                 onnx_path: Optional[str] = None, size: Optional[Tuple[int, int]] = None, devices: Sequence[int] = (0,),
                 max_batch: int = 64, max_delay: float = 0.002, capacity: int = 0, workers: int = 2, cap: int = 64,
                 max_frame_bytes: int = 0, backend: Optional[Callable] = None, flags: int = 0, host_chunk: int = 0,
                 annotate_quality: int = 0, annotate_scale: Tuple[float, float] = (0.0, 0.0), annotate_max_bytes: int = 0):
        """backend: test seam — a Python callable (device, frames: list of HxWx3 arrays) -> list of [n,5] arrays that stands
        in for the batched GPU call (`uf_batcher_create_ex`); the product passes None and gets one handle per device."""
        lib = _capi.load()
        wh = size or variant.width_height()
        cfg = _capi.uf_batcher_config()
        cfg.struct_size = C.sizeof(_capi.uf_batcher_config)
        m = cfg.model
        m.struct_size = C.sizeof(_capi.uf_config)
        self._path = os.fsencode(onnx_path or default_model_path(variant))
        m.onnx_path = self._path
        m.net_w, m.net_h = wh
        m.max_iou, m.min_confidence = max_iou, min_confidence
        m.flags, m.host_chunk = flags, host_chunk
        self._devs = (C.c_int32 * len(devices))(*devices)
        cfg.devices, cfg.n_devices = self._devs, len(devices)
        cfg.max_batch, cfg.max_delay_us = max_batch, max(1, int(max_delay * 1e6))
        cfg.capacity, cfg.workers, cfg.det_cap, cfg.max_frame_bytes = capacity, workers, cap, max_frame_bytes
        cfg.annotate_quality, cfg.annotate_max_bytes = annotate_quality, annotate_max_bytes
        cfg.annotate_scale_w, cfg.annotate_scale_h = annotate_scale
        self.file_stride = annotate_max_bytes or (1 << 20)
        self.cap, self.devices = cap, list(devices)
        self._cb = None
        h = C.c_void_p()
        if backend is not None:
            def thunk(user, device, rgb, w, hh, n, out, capn, n_out):
                try:
                    frames = [np.ctypeslib.as_array(C.cast(rgb[i], C.POINTER(C.c_uint8)), (hh[i], w[i], 3)) for i in range(n)]
                    res = backend(int(device), frames)
                    for i, d in enumerate(res):
                        d = np.asarray(d, np.float32).reshape(-1, 5)
                        n_out[i] = len(d)
                        for j, row in enumerate(d[:capn]):
                            o = out[i * capn + j]
                            o.x0, o.y0, o.x1, o.y1, o.conf = (float(v) for v in row)
                    return 0
                except Exception:
                    return 5
            self._cb = _capi.uf_batch_fn(thunk)
            _check(lib.uf_batcher_create_ex(C.byref(cfg), self._cb, None, C.byref(h)))
        else:
            _check(lib.uf_batcher_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._callbacks: Dict[int, Callable] = {}
        self._cb_lock = threading.Lock()
        self._next_tag = 1 << 62
        self._dispatcher: Optional[threading.Thread] = None
        self._closing = False

    # -- producer side (router.rs:64-72)
    def try_submit(self, stream: int, frame: np.ndarray, on_result: Optional[Callable[[int, list], None]] = None,
                   tag: Optional[int] = None) -> bool:
        img = _as_rgb(frame)
        if on_result is not None:
            with self._cb_lock:
                tag = self._next_tag
                self._next_tag += 1
                self._callbacks[tag] = on_result
                if self._dispatcher is None:
                    self._dispatcher = threading.Thread(target=self._dispatch, daemon=True)
                    self._dispatcher.start()
        ok = C.c_int32()
        _check(_capi.load().uf_batcher_try_submit(self._h, stream, img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0],
                                                  tag or 0, C.byref(ok)))
        if not ok.value and on_result is not None:
            with self._cb_lock:
                self._callbacks.pop(tag, None)
        return bool(ok.value)

    def try_submit_jpeg(self, stream: int, jpeg: bytes, tag: int = 0) -> bool:
        """A frame as a baseline JPEG file (N2): decoded on the way (Huffman on the host, the rest on the GPU)."""
        ok = C.c_int32()
        buf = C.create_string_buffer(bytes(jpeg), len(jpeg))
        _check(_capi.load().uf_batcher_try_submit_jpeg(self._h, stream, buf, len(jpeg), tag, C.byref(ok)))
        return bool(ok.value)

    def ingest(self, msg: bytes, tag: int = 0) -> Tuple[bool, int]:
        """One length-delimited data-socket frame (bincode ProtoMsg): parse, key the stream with hashed(&id), queue the JPEG
        payload on the owner GPU. Returns (accepted, stream key)."""
        ok, key = C.c_int32(), C.c_uint64()
        buf = C.create_string_buffer(bytes(msg), len(msg))
        _check(_capi.load().uf_batcher_ingest(self._h, buf, len(msg), tag, C.byref(ok), C.byref(key)))
        return bool(ok.value), int(key.value)

    def acquire(self, stream: int, h: int, w: int):
        """Zero-copy producer (N4): a [h,w,3] view of a pinned slot in the owner GPU's pool + its ticket, or (None, 0) if dropped."""
        buf, ticket = C.c_void_p(), C.c_uint64()
        _check(_capi.load().uf_batcher_acquire(self._h, stream, h * w * 3, C.byref(buf), C.byref(ticket)))
        if not buf.value:
            return None, 0
        return np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), (h, w, 3)), int(ticket.value)

    def commit(self, ticket: int, h: int, w: int, tag: int = 0) -> None:
        _check(_capi.load().uf_batcher_commit(self._h, ticket, w, h, tag))

    def abort(self, ticket: int) -> None:
        _check(_capi.load().uf_batcher_abort(self._h, ticket))

    # -- consumer side (inferer.rs:41-46)
    def poll(self, max_results: int = 256, timeout: float = 0.0) -> List[dict]:
        res = (_capi.uf_result * max_results)()
        dets = np.zeros((max_results, self.cap, 5), np.float32)
        n = C.c_uint32()
        _check(_capi.load().uf_batcher_poll(self._h, res, dets.ctypes.data_as(C.POINTER(_capi.uf_det)), max_results,
                                            int(timeout * 1e3), C.byref(n)))
        return [dict(stream=int(r.stream), tag=int(r.user_tag), device=int(r.device), status=int(r.status), n_dets=int(r.n_dets),
                     batch_size=int(r.batch_size), latency_us=int(r.latency_us), dets=dets[i, : min(r.n_dets, self.cap)].copy())
                for i, r in enumerate(res[: n.value])]

    def poll_frames(self, max_results: int = 64, timeout: float = 0.0) -> List[dict]:
        """Annotate mode (uf_batcher_poll_frames): poll() plus `file` = the frame's annotated JPEG (b"" for RGB-submitted frames)."""
        res = (_capi.uf_result * max_results)()
        dets = np.zeros((max_results, self.cap, 5), np.float32)
        files = np.empty(max_results * self.file_stride, np.uint8)
        n = C.c_uint32()
        _check(_capi.load().uf_batcher_poll_frames(self._h, res, dets.ctypes.data_as(C.POINTER(_capi.uf_det)), files.ctypes.data_as(C.c_void_p),
                                                   self.file_stride, max_results, int(timeout * 1e3), C.byref(n)))
        return [dict(stream=int(r.stream), tag=int(r.user_tag), device=int(r.device), status=int(r.status), n_dets=int(r.n_dets),
                     batch_size=int(r.batch_size), latency_us=int(r.latency_us), dets=dets[i, : min(r.n_dets, self.cap)].copy(),
                     file=files[i * self.file_stride: i * self.file_stride + r.file_bytes].tobytes())
                for i, r in enumerate(res[: n.value])]

    def _dispatch(self) -> None:
        while True:
            got = self.poll(256, 0.05)
            if not got and self._closing:
                return
            for r in got:
                with self._cb_lock:
                    cb = self._callbacks.pop(r["tag"], None)
                if cb is not None and r["status"] == 0:  # a failed batch skips its frames (`if let Ok(..)`, inferer.rs:37)
                    cb(r["stream"], [((float(d[0]), float(d[1]), float(d[2]), float(d[3])), float(d[4])) for d in r["dets"]])

    def drive(self, frames_ptr: int, n_frames: int, w: int, h: int, streams: Sequence[int], total: int, producers: int = 4):
        """Measurement aid (uf_debug_batcher_drive): C++ producer threads + C++ poll loop; returns (seconds, detections)."""
        arr = (C.c_uint64 * len(streams))(*streams)
        sec, det = C.c_double(), C.c_uint64()
        _check(_capi.load().uf_debug_batcher_drive(self._h, C.c_void_p(frames_ptr), n_frames, w, h, arr, len(streams), total, producers,
                                                   C.byref(sec), C.byref(det)))
        return float(sec.value), int(det.value)

    def drive_msgs(self, msgs: Sequence[bytes], total: int, producers: int = 4):
        """Measurement aid (uf_debug_batcher_drive_msgs): C++ producers hand wire messages (bincode ProtoMsg::FrameMsg with a JPEG
        payload) to uf_batcher_ingest, a C++ loop polls; returns (seconds, detections)."""
        arr = (C.c_char_p * len(msgs))(*msgs)
        lens = (C.c_size_t * len(msgs))(*[len(x) for x in msgs])
        sec, det = C.c_double(), C.c_uint64()
        _check(_capi.load().uf_debug_batcher_drive_msgs(self._h, arr, lens, len(msgs), total, producers, C.byref(sec), C.byref(det)))
        return float(sec.value), int(det.value)

    def flush(self, timeout: float = 30.0) -> None:
        _check(_capi.load().uf_batcher_flush(self._h, int(timeout * 1e3)))

    def stats(self) -> dict:
        if not self._h.value:
            return dict(self._final_stats)  # closed: the counters as they stood
        s = _capi.uf_batcher_stats()
        _check(_capi.load().uf_batcher_stats_read(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in s._fields_}

    def owner(self, stream: int) -> int:
        d = C.c_int32()
        _check(_capi.load().uf_batcher_owner(self._h, stream, C.byref(d)))
        return int(d.value)

    def model_handle(self, device_slot: int = 0) -> Any:
        """Borrowed `uf_model*` of one device (parity hooks / profiling on the batcher's own handle)."""
        h = C.c_void_p()
        _check(_capi.load().uf_batcher_model(self._h, device_slot, C.byref(h)))
        return h

    def close(self) -> None:
        """Finish what is queued, deliver it, join the workers, free the handles."""
        if getattr(self, "_h", None) is None or not self._h.value:
            return
        try:
            self.flush()
        except UltrafaceError:
            pass
        self._closing = True
        if self._dispatcher is not None:
            self._dispatcher.join()
        self._final_stats = self.stats()
        _capi.load().uf_batcher_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
