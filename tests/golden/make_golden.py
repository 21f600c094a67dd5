"""Generates tests/golden/*.npz from the reference's own test photos (run HERE, where
/root/reference exists; the GPU box has no /root/reference, so the vectors are committed).

- test_pics/*.jpg: the eight `resources/test_pics` JPEGs themselves (640x427 x5, 640x462, 640x676,
  640x960; integration_tests.rs:20-29), copied byte for byte (460 KB) so the tests run the
  reference's own geometry (640 -> 640 identity / 640 -> 320 horizontally, non-dyadic vertically).
  They are decoded with PIL to RGB8 at test time. NOTE: PIL (libjpeg-turbo) pixels can differ by a
  few LSB from the reference's `jpeg-decoder 0.3.0`; decode is *before* the hot path.
- oracle_pins.npz: outputs of the CPU oracle on those inputs (sha256 of the resized bytes at both
  network sizes, a slice of raw scores/boxes and the detection list for the seed-0 RFB-320
  fixture model) so that an accidental change of oracle or fixture shows up on CPU.
"""
import hashlib
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools.onnx_fixture import build_ultraface_onnx  # noqa: E402
from oracle import hotpath  # noqa: E402
from oracle.ultraface_ref import UltrafaceOracle  # noqa: E402

PICS = "/root/reference/resources/test_pics"
HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = sorted(os.listdir(PICS))


def main():
    pics = {}
    for n in NAMES:
        src, dst = os.path.join(PICS, n), os.path.join(HERE, "test_pics", n)
        if not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            open(dst, "wb").write(open(src, "rb").read())
        pics[n.split("-unsplash")[0]] = np.ascontiguousarray(np.asarray(Image.open(dst).convert("RGB")))

    pins = {}
    model = UltrafaceOracle(build_ultraface_onnx(320, 240, seed=0), 320, 240, 0.5, 0.5)
    for k, im in pics.items():
        for (w, h) in ((320, 240), (640, 480)):
            r = hotpath.resize_triangle(im, w, h)
            pins[f"{k}/resize{w}x{h}/sha256"] = np.frombuffer(hashlib.sha256(r.tobytes()).digest(), np.uint8)
            pins[f"{k}/resize{w}x{h}/row7"] = r[7].copy()
        s, b = model.raw([im])
        pins[f"{k}/scores_head"] = s[0, :32].copy()
        pins[f"{k}/boxes_head"] = b[0, :32].copy()
        dets, idx = hotpath.postproc(s[0], b[0], 0.5, 0.5)
        pins[f"{k}/det_count"] = np.asarray([len(dets)])
        pins[f"{k}/det_idx_head"] = idx[:16].copy()
    np.savez_compressed(os.path.join(HERE, "oracle_pins.npz"), **pins)
    for f in ("oracle_pins.npz",):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
