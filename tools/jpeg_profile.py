"""Where the time of the JPEG-input path goes: per-kernel-family device times (event pairs) and host wall-clock of the
host-side preparation, for the bench's JPEG workload. Run on a GPU box: python tools/jpeg_profile.py [batch]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from infercam_onnx_b200 import nn  # noqa: E402


def main():
    import cv2
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    tmp = tempfile.TemporaryDirectory()
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    path, w, h = bench.make_model_file(tmp.name, args)
    model = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, size=(w, h), max_batch=B)
    src = bench.smooth_frames(32, seed=7)
    files = []
    for f in src:
        ok, buf = cv2.imencode(".jpg", f[:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                       cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
        files.append(buf.tobytes())
    jpegs = [files[i % len(files)] for i in range(B)]
    for _ in range(3):
        model.run_batch_jpeg(jpegs, cap=64)
    t0 = time.perf_counter()
    for _ in range(10):
        model.run_batch_jpeg(jpegs, cap=64)
    dt = (time.perf_counter() - t0) / 10
    print(f"batch {B}: {dt * 1e3:.3f} ms per call, {B / dt:.0f} frames/s (wall clock, pipelined)")
    model.profile_enable(True)
    model.profile_reset()
    for _ in range(5):
        model.run_batch_jpeg(jpegs, cap=64)
    rows = model.profile_by_family()
    tot = sum(r["device_ms"] for r in rows)
    for r in sorted(rows, key=lambda r: -r["device_ms"]):
        print(f"  {r['name']:<34} {r['device_ms'] / 5:8.3f} ms/call  {100 * r['device_ms'] / tot:5.1f}%  launches/call {r['launches'] / 5:.0f}")
    print(f"  device total {tot / 5:.3f} ms/call")
    _, launches = model.jpeg_coefficients_gpu(files[0])
    print("sync launches for one frame:", launches, "bytes", len(files[0]))
    model.close()


if __name__ == "__main__":
    main()
