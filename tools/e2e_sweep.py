import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from infercam_onnx_b200 import nn
from tools.onnx_fixture import write_ultraface_onnx
tmp = tempfile.mkdtemp()
path = write_ultraface_onnx(os.path.join(tmp, "m.onnx"), width=320, height=240, seed=0, cls_bias=-0.75)
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 4
hc = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=256, slots=slots, host_chunk=hc)
frames = np.random.default_rng(0).integers(0, 256, (256, 480, 640, 3), dtype=np.uint8)
pin = nn.PinnedFrames(256, 480, 640); pin.array[:] = frames
for _ in range(5): m.run_batch_ptr(pin.ptr, 640, 480, 256, cap=128)
t = time.perf_counter()
for _ in range(20): m.run_batch_ptr(pin.ptr, 640, 480, 256, cap=128)
dt = (time.perf_counter() - t) / 20
print(f"host_chunk={hc or 'auto'} slots={slots}: {dt*1e3:.2f} ms/batch {256/dt:.0f} fps", flush=True)
