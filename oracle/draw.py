"""CPU restatement of the rectangle overlay of `draw_bboxes_on_image` (infer_server/src/inferer.rs:58-92): the reference's
corner arithmetic and casts, and imageproc's `draw_hollow_rect` (crate imageproc, not vendored: four line segments
(left, top)-(right, top), (left, bottom)-(right, bottom), (left, top)-(left, bottom), (right, top)-(right, bottom) with
right = left + width - 1, bottom = top + height - 1, every point outside the image skipped). The text overlay (`draw_text`,
rusttype) is not restated. TEST INFRASTRUCTURE; PARITY STATUS: unpinned (no Rust toolchain, crate absent)."""
from __future__ import annotations

import numpy as np


def _as_i32(v: np.float32) -> int:  # Rust `as i32`: truncation, saturating, NaN -> 0
    if np.isnan(v):
        return 0
    return int(max(-2**31, min(2**31 - 1, int(np.trunc(np.float64(v)))))) if np.isfinite(v) else (2**31 - 1 if v > 0 else -2**31)


def _as_u32(v: np.float32) -> int:
    if np.isnan(v) or v <= 0:
        return 0
    return int(min(2**32 - 1, int(np.trunc(np.float64(v))))) if np.isfinite(v) else 2**32 - 1


def draw_boxes(rgb: np.ndarray, dets, width: float, height: float) -> np.ndarray:
    out = np.array(rgb, copy=True)
    H, W = out.shape[:2]
    width, height = np.float32(width), np.float32(height)
    for d in np.asarray(dets, np.float32).reshape(-1, 5):
        x_tl, y_tl = d[0] * width, d[1] * height
        x_br, y_br = d[2] * width, d[3] * height
        rw, rh = _as_u32(x_br - x_tl), _as_u32(y_br - y_tl)
        if rw == 0 or rh == 0:
            continue  # Rect::of_size would panic
        left, top = _as_i32(x_tl), _as_i32(y_tl)
        right, bottom = left + rw - 1, top + rh - 1
        xs = np.arange(max(left, 0), min(right, W - 1) + 1)
        ys = np.arange(max(top, 0), min(bottom, H - 1) + 1)
        for y in (top, bottom):
            if 0 <= y < H:
                out[y, xs] = (0, 255, 0)
        for x in (left, right):
            if 0 <= x < W:
                out[ys, x] = (0, 255, 0)
    return out
