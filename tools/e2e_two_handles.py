"""Experiment: two model handles fed from two host threads (two batches in flight) vs one."""
import os, sys, time, tempfile, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from infercam_onnx_b200 import nn
from tools.onnx_fixture import write_ultraface_onnx
tmp = tempfile.mkdtemp()
path = write_ultraface_onnx(os.path.join(tmp, "m.onnx"), width=320, height=240, seed=0, cls_bias=-0.75)
nh = int(sys.argv[1]) if len(sys.argv) > 1 else 2
models = [nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=256) for _ in range(nh)]
frames = np.random.default_rng(0).integers(0, 256, (256, 480, 640, 3), dtype=np.uint8)
pins = []
for _ in range(nh):
    p = nn.PinnedFrames(256, 480, 640); p.array[:] = frames; pins.append(p)
steps = 20
def work(i, n):
    for _ in range(n): models[i].run_batch_ptr(pins[i].ptr, 640, 480, 256, cap=128)
for i in range(nh): work(i, 3)
t = time.perf_counter()
ths = [threading.Thread(target=work, args=(i, steps)) for i in range(nh)]
[x.start() for x in ths]; [x.join() for x in ths]
dt = time.perf_counter() - t
print(f"handles={nh}: {nh*steps*256/dt:.0f} fps  ({dt/steps/nh*1e3:.2f} ms per batch)")
