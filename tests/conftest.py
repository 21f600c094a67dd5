import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def onnx_dir(tmp_path_factory):
    return tmp_path_factory.mktemp("onnx")


@pytest.fixture(scope="session")
def make_onnx(onnx_dir):
    """Writes (and caches) a seeded random-init UltraFace ONNX file; returns its path."""
    from tools.onnx_fixture import write_ultraface_onnx
    cache = {}

    def _make(width=320, height=240, variant="RFB", seed=0, with_bn=False, cls_bias=0.0, head_gain=1.0, style="simplified",
              tail="standard"):
        key = (width, height, variant, seed, with_bn, cls_bias, head_gain, style, tail)
        if key not in cache:
            path = os.path.join(str(onnx_dir), "uf_%d_%d_%s_%d_%d_%g_%g_%s_%s.onnx" % key)
            write_ultraface_onnx(path, width=width, height=height, variant=variant, seed=seed, with_bn=with_bn,
                                 cls_bias=cls_bias, head_gain=head_gain, style=style, tail=tail)
            cache[key] = path
        return cache[key]

    return _make


@pytest.fixture(scope="session")
def test_pics():
    """The reference's eight resources/test_pics at full size (640 px wide), decoded with PIL to RGB8; key = file name up
    to "-unsplash" (tests/golden/make_golden.py)."""
    from PIL import Image
    d = os.path.join(ROOT, "tests", "golden", "test_pics")
    return {n.split("-unsplash")[0]: np.ascontiguousarray(np.asarray(Image.open(os.path.join(d, n)).convert("RGB")))
            for n in sorted(os.listdir(d)) if n.endswith(".jpg")}


# integration_tests.rs:20-29: faces per photo, RFB-640 at 0.5 / 0.5 with the downloaded weights
REFERENCE_FACE_COUNTS = {"bruce-mars-ZXq7xoo98b0": 3, "clarke-sanders-ybPJ47PMT_M": 6, "helena-lopes-e3OUQGT9bWU": 4,
                         "kaleidico-d6rTXEtOclk": 3, "michael-dam-mEZ3PoFGs_k": 1, "mika-W0i1N6FdCWA": 1,
                         "omar-lopez-T6zu4jFhVwg": 10, "ken-cheung-KonWFWUaAuk": 0}


@pytest.fixture(scope="session")
def reference_face_counts():
    return dict(REFERENCE_FACE_COUNTS)


@pytest.fixture(scope="session")
def real_rfb640_path():
    """The weight file the reference downloads (nn.rs:21, 145-157); absent here (no network) -> the tests that need it skip."""
    cache = os.environ.get("XDG_CACHE_HOME") or os.path.join(os.path.expanduser("~"), ".cache")
    for p in (os.environ.get("ULTRAFACE_RFB640_ONNX", ""), os.path.join(cache, "infercam_onnx", "ultraface-RFB-640.onnx")):
        if p and os.path.exists(p):
            return p
    pytest.skip("ultraface-RFB-640.onnx not present (~/.cache/infercam_onnx/, or $ULTRAFACE_RFB640_ONNX): the reference downloads it")


@pytest.fixture(scope="session")
def oracle_pins():
    z = np.load(os.path.join(ROOT, "tests", "golden", "oracle_pins.npz"))
    return {k: z[k] for k in z.files}
