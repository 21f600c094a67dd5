import os, sys, time, tempfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from infercam_onnx_b200 import nn, _capi
from tools.onnx_fixture import write_ultraface_onnx
tmp = tempfile.mkdtemp()
path = write_ultraface_onnx(os.path.join(tmp, "m.onnx"), width=320, height=240, seed=0, cls_bias=-0.75)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
frames = np.random.default_rng(0).integers(0, 256, (256, 480, 640, 3), dtype=np.uint8)
m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=256, flags=flags)
d = torch.from_numpy(frames).cuda()
for _ in range(5): m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(30): m.run_batch_device(d.data_ptr(), 640, 480, 256, cap=128)
dt = (time.perf_counter() - t) / 30
pin = nn.PinnedFrames(1, 480, 640); pin.array[:] = frames[:1]
lat = []
for i in range(320):
    t = time.perf_counter(); m.run_batch_ptr(pin.ptr, 640, 480, 1, cap=128); lat.append((time.perf_counter() - t) * 1e3)
lat = sorted(lat[20:])
print(f"flags={flags}: value {256/dt:.0f} fps ({dt*1e3:.3f} ms)  b1 p50 {lat[len(lat)//2]:.3f} ms")
