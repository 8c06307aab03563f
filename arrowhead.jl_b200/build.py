"""In-tree build of libarrowhead_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
SOURCES = [os.path.join(_PKG, "csrc", "ah_api.cu")]
HEADERS = [os.path.join(_PKG, "csrc", f) for f in ("ah_device.cuh", "ah_full.cuh", "ah_pipe.cuh", "ah_roots.cuh", "ah_gemm.cuh", "ah_multi.inl")] + [
    os.path.join(os.path.dirname(_PKG), "include", "arrowhead_b200.h")]
OUT = os.path.join(_PKG, "libarrowhead_b200.so")

# -fmad=false: the only FMAs in the binary are the explicit ones (secular term, two-product), so the
# rest rounds like the reference's Julia code.  -lineinfo: ncu source pages map to these files.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-ldl"]


def build_library(force: bool = False, verbose: bool = False) -> str:
    newest = max(os.path.getmtime(p) for p in SOURCES + HEADERS)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libarrowhead_b200.so")
    tmp = OUT + f".tmp{os.getpid()}"
    cmd = [nvcc, *NVCC_FLAGS, "-o", tmp, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)     # the image's $CC (/opt/gcc) is not the nvcc host compiler
    subprocess.run(cmd, check=True, env=env)
    os.replace(tmp, OUT)
    return OUT
