"""Builds libultraface_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libultraface_b200.so")
SOURCES = ["engine.cu", "kernels_preproc.cu", "kernels_prestem.cu", "kernels_jpeg.cu", "kernels_jpeg_enc.cu", "kernels_jpeg_henc.cu", "kernels_jpeg_huff.cu", "jpeg_encode.cc", "kernels_conv.cu", "kernels_post.cu", "kernels_tc.cu", "onnx_graph.cc", "plan.cc", "tail_check.cc", "batcher.cc", "jpeg_entropy.cc",
           "resize_taps.cc"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall", "-Xptxas", "-v",
              "-Xlinker", "--no-undefined", "-cudart", "shared"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ultraface_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}); see {log}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
