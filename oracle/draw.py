"""CPU restatement of the rectangle overlay of `draw_bboxes_on_image` (infer_server/src/inferer.rs:58-92): the reference's
corner arithmetic and casts, and imageproc's `draw_hollow_rect` (crate imageproc, not vendored: four line segments
(left, top)-(right, top), (left, bottom)-(right, bottom), (left, top)-(left, bottom), (right, top)-(right, bottom) with
right = left + width - 1, bottom = top + height - 1, every point outside the image skipped), and of the text overlay
(`draw_text`, inferer.rs:80-88): the "{:.2}%" text of confidence * 100 (f32), and imageproc's `draw_text_mut` blend —
per glyph pixel weighted_sum(pixel, colour, 1 - v, v) = clamp(pixel * (1 - v) + colour * v) in f32, truncated — over glyph
coverage that is an INPUT here (the atlas rusttype renders; rasterising is not restated). Detections are drawn in order,
rectangle then text. TEST INFRASTRUCTURE; PARITY STATUS: unpinned (no Rust toolchain, crates absent)."""
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


def confidence_text(conf) -> str:
    """format!("{:.2}%", confidence * 100.0) with f32 arithmetic: the f32 product's exact value, two decimals."""
    pct = np.float32(conf) * np.float32(100.0)
    return "%.2f%%" % float(pct)


def _blend(pix: int, col: float, v: np.float32) -> int:
    """imageproc weighted_channel_sum + Clamp<f32> for u8."""
    lw = np.float32(1.0) - v
    x = np.float32(np.float32(pix) * lw) + np.float32(np.float32(col) * v)
    x = np.float32(x)
    if x < np.float32(255.0):
        return int(x) if x > np.float32(0.0) else 0
    return 255


def draw_text(out: np.ndarray, x: int, y: int, text: str, atlas) -> None:
    """atlas = (charset, max_len, glyphs[pos][k] = (x0, y0, w, h, offset), coverage f32[]). In place."""
    charset, max_len, glyphs, cov = atlas
    H, W = out.shape[:2]
    colour = (0.0, 255.0, 0.0)
    for pos, ch in enumerate(text):
        k = charset.find(ch)
        if pos >= max_len or k < 0:
            continue
        x0, y0, w, h, off = (int(t) for t in glyphs[pos][k])
        for gy in range(h):
            for gx in range(w):
                ix, iy = x + x0 + gx, y + y0 + gy
                if 0 <= ix < W and 0 <= iy < H:
                    v = np.float32(cov[off + gy * w + gx])
                    for c in range(3):
                        out[iy, ix, c] = _blend(int(out[iy, ix, c]), colour[c], v)


def draw_boxes(rgb: np.ndarray, dets, width: float, height: float, atlas=None) -> np.ndarray:
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
        if atlas is not None:
            draw_text(out, left, top, confidence_text(d[4]), atlas)
    return out
