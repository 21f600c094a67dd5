import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from infercam_onnx_b200 import nn, _capi
from tools.onnx_fixture import write_ultraface_onnx
tmp = tempfile.mkdtemp()
path = write_ultraface_onnx(os.path.join(tmp, "m.onnx"), width=320, height=240, seed=0, cls_bias=-0.75)
frames = np.random.default_rng(0).integers(0, 256, (1, 480, 640, 3), dtype=np.uint8)
for flags, label in [(0, "default"), (_capi.UF_FLAG_NO_TC, "no_tc"), (_capi.UF_FLAG_NO_TC | _capi.UF_FLAG_NO_GRAPH, "no_tc,no_graph"), (_capi.UF_FLAG_NO_GRAPH, "no_graph")]:
    m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, max_batch=1, flags=flags)
    pin = nn.PinnedFrames(1, 480, 640); pin.array[:] = frames
    lat = []
    for i in range(320):
        t = time.perf_counter(); m.run_batch_ptr(pin.ptr, 640, 480, 1, cap=128); lat.append((time.perf_counter() - t) * 1e3)
    lat = sorted(lat[20:])
    print(f"{label:16s} p50 {lat[len(lat)//2]:.3f} ms  p99 {lat[int(len(lat)*0.99)]:.3f} ms  launches/frame {m.launch_count()//320}")
    m.close()
