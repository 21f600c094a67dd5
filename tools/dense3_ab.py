"""A/B of the dense 3x3 kernels of the RFB branches: SIMT (default) vs tcgen05 zero-copy implicit GEMM (UF_FLAG_DENSE3_TC)."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from infercam_onnx_b200 import _capi, nn  # noqa: E402
from tools.onnx_fixture import write_ultraface_onnx  # noqa: E402

tmp = tempfile.mkdtemp()
path = write_ultraface_onnx(os.path.join(tmp, "m.onnx"), width=320, height=240, seed=0, cls_bias=-0.75)
frames = np.random.default_rng(0).integers(0, 256, (256, 480, 640, 3), dtype=np.uint8)
d = torch.from_numpy(frames).cuda()
for flags, label in [(0, "simt"), (_capi.UF_FLAG_DENSE3_TC, "tcgen05")]:
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=256, flags=flags)
    for _ in range(3):
        m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
    dt = (time.perf_counter() - t0) / 20
    m.profile_enable(True)
    m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
    m.profile_reset()
    for _ in range(5):
        m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
    print(label, f"{256 / dt:.0f} frames/s")
    for s in sorted(m.profile_read(), key=lambda s: -s["device_ms"]):
        if "dense3x3" in s["name"]:
            print(f"   {s['name']:48s} {s['device_ms'] / 5 * 1e3 / s['launches'] * 5:8.1f} us/launch x {s['launches'] // 5}")
    m.profile_enable(False)
    m.close()
