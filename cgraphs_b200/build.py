"""Build libhbgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libhbgpu.so")
LIB_DEBUG = os.path.join(CSRC, "libhbgpu_debug.so")      # -DHB_DEBUG: bounds asserts on scratch indices (HBGPU_LIB selects it)
SOURCES = [os.path.join(CSRC, "hbgpu.cu")]
HEADERS = [os.path.join(HERE, "..", "include", "hbgpu.h")] + \
          [os.path.join(CSRC, f) for f in ("hb_common.cuh", "hb_fused.cuh", "hb_general.cuh", "hb_finalize.cuh", "hb_wires.cuh", "hb_framepairs.cuh", "hb_xtc.cuh", "hb_xtc_host.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared"]


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS if os.path.exists(f))


def build(force=False, verbose=False, debug=False):
    lib = LIB_DEBUG if debug else LIB
    if not force and not needs_build(lib):
        return lib
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-DHB_DEBUG"] if debug else []) + (["-Xptxas", "-v"] if verbose else []) + ["-o", lib] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        sys.stderr.write(res.stderr)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
