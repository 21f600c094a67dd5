"""GPU bring-up diagnostics (development aid, not a test): prints instead of asserting so one
gpurun call shows every mismatch. Usage: python tools/gpu_diag.py [--quick]"""
import os
import sys
import tempfile
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from infercam_onnx_b200 import _capi, nn  # noqa: E402
from tools.onnx_fixture import write_ultraface_onnx  # noqa: E402
from oracle import hotpath  # noqa: E402
from oracle.ultraface_ref import UltrafaceOracle  # noqa: E402

quick = "--quick" in sys.argv
tmp = tempfile.mkdtemp()
path = write_ultraface_onnx(os.path.join(tmp, "m.onnx"), width=320, height=240, seed=0, cls_bias=-0.75)
oracle = UltrafaceOracle(path, 320, 240, 0.5, 0.5)
rng = np.random.default_rng(0)


def section(t):
    print("\n==== " + t, flush=True)


def guard(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()


def resize_checks():
    section("resize")
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path)
    for shape in [(480, 640), (427, 640), (720, 1280), (240, 320), (100, 100), (7, 9)]:
        im = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
        got, ref = m.preproc_u8(im), hotpath.resize_triangle(im, 320, 240)
        d = np.abs(got.astype(int) - ref.astype(int))
        print(shape, "mismatch px:", int((d > 0).sum()), "max:", int(d.max()))
    im = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    exp = hotpath.normalise_nchw(hotpath.resize_triangle(im, 320, 240), 0)[None]
    print("normalise equal:", np.array_equal(m.preproc(im), exp))
    m.close()


def layer_checks():
    import torch
    frame = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    with torch.no_grad():
        (s_ref, b_ref), env = oracle.net(torch.from_numpy(oracle.preproc(frame)), keep=True)
    for flags, label in [(_capi.UF_FLAG_FORCE_GENERIC, "generic"), (_capi.UF_FLAG_NO_FUSION, "unfused"), (0, "fused")]:
        section("layers: " + label)
        try:
            m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, flags=flags)
            got = m.run(frame, cap=2048)
            for name, (idx, c, h, w) in m.tensors().items():
                if name not in env:
                    continue
                t = m.tensor_read(idx, 0, (c, h, w))
                ref = env[name].numpy()[0]
                d = float(np.abs(t - ref).max())
                flag = "" if d <= 1e-4 * max(1.0, float(np.abs(ref).max())) else "   <<<<<< BAD"
                print(f"{name:12s} {str((c, h, w)):16s} maxdiff {d:.3e} refmax {float(np.abs(ref).max()):.3f}{flag}")
            s, b = m.raw_outputs(0, 1)
            print("raw scores maxdiff", float(np.abs(s - s_ref.numpy()).max()), "boxes", float(np.abs(b - b_ref.numpy()).max()))
            ref, _ = hotpath.postproc(s[0], b[0], 0.5, 0.5)
            print("dets gpu", len(got), "oracle(on gpu raw)", len(ref), "equal:",
                  len(got) == len(ref) and np.array_equal(np.float32([[*bb, c] for bb, c in got]).reshape(-1, 5), ref))
            m.close()
        except Exception:
            traceback.print_exc()


def post_checks():
    section("postproc")
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path)
    for K, spread in [(5, 0.3), (300, 0.05), (4420, 0.3), (4420, 0.02), (17640, 0.2), (20000, 0.01)]:
        r = np.random.default_rng(K)
        c = r.random((K, 2)).astype(np.float32) * 0.6 + 0.2
        wh = (r.random((K, 2)).astype(np.float32) * spread).astype(np.float32)
        boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        scores = r.random((K, 2)).astype(np.float32)
        scores[r.integers(0, K, K // 5), 1] = np.float32(0.75)
        t0 = time.perf_counter()
        dets, idx = m.postproc(scores, boxes)
        t1 = time.perf_counter()
        ref, ridx = hotpath.postproc(scores, boxes, 0.5, 0.5)
        print(K, spread, "n gpu", len(idx), "n ref", len(ridx), "idx equal", np.array_equal(idx, ridx), "dets equal",
              np.array_equal(dets, ref), f"{(t1 - t0) * 1e3:.2f} ms")
    m.close()


def speed():
    import torch
    section("speed (batch 256, device resident)")
    for flags, label in [(_capi.UF_FLAG_NO_FUSION, "unfused"), (0, "fused")]:
        m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=256, flags=flags)
        frames = rng.integers(0, 256, (256, 480, 640, 3), dtype=np.uint8)
        d = torch.from_numpy(frames).cuda()
        for _ in range(3):
            m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            dets, counts = m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
        dt = (time.perf_counter() - t0) / 10
        print(label, f"{dt * 1e3:.3f} ms/batch  {256 / dt:.0f} frames/s  mean dets {np.mean(counts):.1f}")
        m.profile_enable(True)
        m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
        m.profile_reset()
        for _ in range(5):
            m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
        for s in sorted(m.profile_by_family(), key=lambda s: -s["device_ms"]) + [dict(name="-- per layer --", device_ms=1e-9, launches=0, algorithmic_bytes=0, compulsory_bytes=0, flops=0)] + sorted(m.profile_read(), key=lambda s: -s["device_ms"])[:30]:
            ms = s["device_ms"] / 5
            print(f"   {s['name']:48s} {ms:8.3f} ms/batch  launches {s['launches'] // 5:4d}  alg {s['algorithmic_bytes'] / 5 / ms / 1e6:8.1f} GB/s"
                  f"  min {s['compulsory_bytes'] / 5 / ms / 1e6:8.1f} GB/s  {s['flops'] / 5 / ms / 1e9:7.2f} TFLOP/s")
        m.profile_enable(False)
        m.close()


guard(resize_checks)
guard(layer_checks)
guard(post_checks)
if not quick:
    guard(speed)
print("\ndiag done", flush=True)
