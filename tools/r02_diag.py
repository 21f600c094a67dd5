"""Round-2 diagnostics (development aid): error of the tensor-core (3xTF32) and SIMT fp32 paths against the fp32 and fp64
oracle on the BN-explicit fixture with real photos (box magnitudes ~7 with random-init weights)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from PIL import Image  # noqa: E402

from infercam_onnx_b200 import _capi, nn  # noqa: E402
from oracle.ultraface_ref import UltrafaceOracle  # noqa: E402
from tools.onnx_fixture import write_ultraface_onnx  # noqa: E402

d = os.path.join(ROOT, "tests", "golden", "test_pics")
pics = [np.ascontiguousarray(np.asarray(Image.open(os.path.join(d, n)).convert("RGB"))) for n in sorted(os.listdir(d))]
rng = np.random.default_rng(11)
frames = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)] + pics
for tag, kw in (("bn seed3", dict(with_bn=True, seed=3)), ("plain seed0", dict(seed=0)), ("rfb640", dict(width=640, height=480, seed=0))):
    w, h = kw.get("width", 320), kw.get("height", 240)
    path = write_ultraface_onnx("/tmp/diag_%s.onnx" % tag.replace(" ", "_"), cls_bias=-0.75, **kw)
    s32, b32 = UltrafaceOracle(path, w, h).raw(frames)
    s64, b64 = UltrafaceOracle(path, w, h, dtype=torch.float64).raw(frames)
    print(f"[{tag}] max|box| {np.abs(b64).max():.2f}  oracle fp32 vs fp64: score {np.abs(s32 - s64).max():.2e} box {np.abs(b32 - b64).max():.2e}")
    for fname, flags in (("default(3xTF32)", 0), ("NO_TC(fp32 SIMT)", _capi.UF_FLAG_NO_TC), ("generic", _capi.UF_FLAG_FORCE_GENERIC)):
        m = nn.UltrafaceModel.new(nn.UltrafaceVariant.W320H240, 0.5, 0.5, onnx_path=path, size=(w, h), max_batch=16, flags=flags)
        m.run_batch(frames, cap=64)
        s, b = m.raw_outputs(0, len(frames))
        m.close()
        rel = np.abs(b - b64) / np.maximum(1.0, np.abs(b64))
        print(f"   {fname:18s} vs fp64: score {np.abs(s - s64).max():.2e} box abs {np.abs(b - b64).max():.2e} box rel(max(1,|ref|)) {rel.max():.2e}"
              f" | vs fp32 oracle: score {np.abs(s - s32).max():.2e} box {np.abs(b - b32).max():.2e}")
