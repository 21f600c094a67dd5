"""Short workload for ncu: N device-resident batches of the bench configuration (RFB-320, 256 frames)."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from infercam_onnx_b200 import nn  # noqa: E402
from tools.onnx_fixture import write_ultraface_onnx  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
net = sys.argv[2] if len(sys.argv) > 2 else "320x240"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 256
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cls_bias = float(sys.argv[5]) if len(sys.argv) > 5 else -0.75
w, h = (int(v) for v in net.split("x"))
tmp = tempfile.mkdtemp()
path = write_ultraface_onnx(os.path.join(tmp, "m.onnx"), width=w, height=h, seed=0, cls_bias=cls_bias)
m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, size=(w, h), max_batch=batch, slots=1, flags=flags)
frames = np.random.default_rng(0).integers(0, 256, (batch, 480, 640, 3), dtype=np.uint8)
d = torch.from_numpy(frames).cuda()
for _ in range(iters):
    dets, counts = m.run_batch_device(d.data_ptr(), 640, 480, batch, cap=128)
print("launches", m.launch_count(), "mean dets", float(np.mean(counts)))
m.close()
