"""A/B of the two fused dw+1x1 kernels of the big maps: SIMT 1x1 (UF_FLAG_TMA_SIMT_PW) vs tcgen05 1x1 (default)."""
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
for flags, label in [(_capi.UF_FLAG_TMA_SIMT_PW, "simt 1x1"), (0, "tcgen05 1x1")]:
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=256, flags=flags)
    for _ in range(3):
        m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
    m.profile_enable(True)
    m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
    m.profile_reset()
    for _ in range(5):
        m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
    print(label)
    for s in sorted(m.profile_read(), key=lambda s: -s["device_ms"]):
        if "tma" in s["name"]:
            print(f"   {s['name']:48s} {s['device_ms'] / 5 / 2 * 1e3:8.1f} us/launch")
    m.profile_enable(False)
    m.close()
