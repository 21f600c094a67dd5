"""The annotate path (N3) for ncu / timing: uf_worker_batch_jpeg over the bench's MJPG workload. Run on a GPU box:
python tools/worker_profile.py [batch]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from infercam_onnx_b200 import nn  # noqa: E402


def main():
    import cv2
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    tmp = tempfile.TemporaryDirectory()
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    path, w, h = bench.make_model_file(tmp.name, args)
    model = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, size=(w, h), max_batch=B, lanes=4)
    files = []
    for f in bench.smooth_frames(32, seed=7):
        ok, buf = cv2.imencode(".jpg", f[:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                       cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
        files.append(buf.tobytes())
    jpegs = [files[i % len(files)] for i in range(B)]
    for _ in range(3):
        out = model.worker_batch_jpeg(jpegs, 1280.0, 720.0, quality=95, cap=64, keep_files=False)
    t0 = time.perf_counter()
    for _ in range(5):
        out = model.worker_batch_jpeg(jpegs, 1280.0, 720.0, quality=95, cap=64, keep_files=False)
    dt = (time.perf_counter() - t0) / 5
    print(f"batch {B}: {dt * 1e3:.3f} ms per call, {B / dt:.0f} frames/s, {sum(out[2]) / B:.0f} bytes per annotated file, {sum(out[1])} detections")
    model.close()


if __name__ == "__main__":
    main()
