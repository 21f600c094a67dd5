"""Small workload touching every round-2 kernel once, for compute-sanitizer (memcheck / racecheck / initcheck):
fused resize+stem (interior and border tiles, both nets' geometries), the NMS bit-matrix path (frames with > 512 candidates,
odd candidate counts), JPEG decode with Huffman decoding on the device and on the host (4:4:4 / 4:2:2 / 4:2:0 / grey, odd sizes,
truncated and bit-flipped files that the device decoder hands back), rectangles + text + JPEG encode (per frame with the host
Huffman coder, and the batch forms with Huffman coding on the device), the batcher."""
import io
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from infercam_onnx_b200 import nn  # noqa: E402
from infercam_onnx_b200.batcher import StreamBatcher  # noqa: E402
from tools.onnx_fixture import write_ultraface_onnx  # noqa: E402

rng = np.random.default_rng(0)
p = write_ultraface_onnx("/tmp/san320.onnx", width=320, height=240, seed=0, cls_bias=-0.75)
ph = write_ultraface_onnx("/tmp/san320h.onnx", width=320, height=240, seed=0, cls_bias=3.0)
m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=p, max_batch=8)
frames = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(3)] + [rng.integers(0, 256, (301, 333, 3), dtype=np.uint8)]
d, c = m.run_batch(frames, cap=64)
m.prestem_u8(np.stack(frames[:3]))


def enc(a, q=90, ss=2):
    b = io.BytesIO()
    Image.fromarray(a).save(b, "JPEG", quality=q, subsampling=ss)
    return b.getvalue()


jpegs = [enc(frames[0], 90, 1), enc(frames[3], 80, 2), enc(frames[3][:9, :17], 85, 0), enc(frames[1], 95, 2)]
g = io.BytesIO()
Image.fromarray(frames[3]).convert("L").save(g, "JPEG", quality=90)
jpegs.append(g.getvalue())
cut = jpegs[0][: len(jpegs[0]) // 2] + b"\xff\xd9"                      # truncated: the device decoder declines, the host redoes it
flip = bytearray(jpegs[3]); flip[len(flip) // 2] ^= 0x10                  # damaged in the middle of the scan
big = enc(np.repeat(np.repeat(frames[0], 2, 0), 2, 1), 92, 1)             # 1280x960: several CTAs per frame
dj, cj = m.run_batch_jpeg(jpegs + [cut, bytes(flip), big], cap=64)
for j in jpegs + [cut, bytes(flip)]:
    m.jpeg_decode_rgb(j)
    m.jpeg_coefficients_gpu(j)
from infercam_onnx_b200 import _capi  # noqa: E402
mhh = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=p, max_batch=8, flags=_capi.UF_FLAG_JPEG_HOST_HUFFMAN)
assert mhh.run_batch_jpeg(jpegs + [cut, bytes(flip), big], cap=64)[1] == cj
mhh.close()
arng = np.random.default_rng(5)
aglyphs, acov, off = [], [], 0
for pos in range(7):
    row = []
    for _ in "0123456789.%":
        gw, gh = int(arng.integers(0, 10)), int(arng.integers(1, 13))
        row.append((pos * 8 + int(arng.integers(-1, 2)), int(arng.integers(0, 5)), gw, gh, off))
        acov.append(arng.random(gw * gh).astype(np.float32))
        off += gw * gh
    aglyphs.append(row)
m.text_atlas_set("0123456789.%", 7, aglyphs, np.concatenate(acov))
boxes = np.float32([[0.1, 0.2, 0.3, 0.6, 0.9], [-0.2, -0.1, 0.25, 0.3, 0.7], [0.8, 0.7, 1.4, 1.3, 0.6], [0, 0, 1, 1, 0.5]])
for f in (frames[0], frames[3], frames[3][:9, :17]):
    m.draw_boxes(f, boxes, float(f.shape[1]), float(f.shape[0]))
    m.annotate_encode_jpeg(np.ascontiguousarray(f), boxes, float(f.shape[1]), float(f.shape[0]), 95)
m.annotate_encode_jpeg(jpegs[1], boxes, 333.0, 301.0, 90)
wd, wc, wf = m.worker_batch_jpeg(jpegs + [cut, bytes(flip)], 640.0, 480.0, quality=95, cap=64)   # decode, detect, draw, encode in one call
rb = m.annotate_reencode_batch_jpeg(jpegs[:3], [boxes, boxes[:1], boxes[:0]], 640.0, 480.0, quality=60)
assert rb[0] == m.annotate_encode_jpeg(jpegs[0], boxes, 640.0, 480.0, 60)
for K, seed in ((700, 1), (4420, 2), (5001, 3)):
    s = rng.random((K, 2)).astype(np.float32)
    c0 = rng.random((K, 2)).astype(np.float32) * 0.6 + 0.2
    wh = rng.random((K, 2)).astype(np.float32) * 0.1
    m.postproc(s, np.concatenate([c0 - wh / 2, c0 + wh / 2], 1))
m.close()
mh = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=ph, max_batch=4)
dh, ch = mh.run_batch(frames[:3], cap=4420)
mh.close()
b = StreamBatcher(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=p, max_batch=4, cap=64)
for i in range(6):
    b.try_submit(i, frames[i % 3], tag=i)
b.try_submit_jpeg(7, jpegs[0], tag=7)
b.flush()
r = b.poll(64)
b.close()
print("sanitize target ok: dets", c, cj, "heavy", ch, "batcher results", len(r))
