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
    from infercam_onnx_b200.onnx_fixture import write_ultraface_onnx
    cache = {}

    def _make(width=320, height=240, variant="RFB", seed=0, with_bn=False, cls_bias=0.0, head_gain=1.0):
        key = (width, height, variant, seed, with_bn, cls_bias, head_gain)
        if key not in cache:
            path = os.path.join(str(onnx_dir), "uf_%d_%d_%s_%d_%d_%g_%g.onnx" % key)
            write_ultraface_onnx(path, width=width, height=height, variant=variant, seed=seed, with_bn=with_bn,
                                 cls_bias=cls_bias, head_gain=head_gain)
            cache[key] = path
        return cache[key]

    return _make


@pytest.fixture(scope="session")
def test_pics():
    """Four of the reference's resources/test_pics (centre strips), see tests/golden/make_golden.py."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "test_pics.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def oracle_pins():
    z = np.load(os.path.join(ROOT, "tests", "golden", "oracle_pins.npz"))
    return {k: z[k] for k in z.files}
