"""TEST INFRASTRUCTURE - builds the C restatements under oracle/ (gcc, no GPU needed).

    python oracle/build_oracle.py

* oracle/libxtc_oracle.so  <- oracle/xtc_codec.c (XTC frame codec restated from the published xdrfile algorithm)

The built library is git-ignored and travels to the GPU box with the snapshot, like the product's own .so.
`__graft_entry__.build()` calls this; building the checker is not using it - only tests/, smoke() and
bench.py's CPU-baseline leg load it (oracle/xtc_port.py).
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
XTC_SRC = os.path.join(HERE, "xtc_codec.c")
XTC_LIB = os.path.join(HERE, "libxtc_oracle.so")


def build(force=False):
    if force or not os.path.exists(XTC_LIB) or os.path.getmtime(XTC_LIB) < os.path.getmtime(XTC_SRC):
        cmd = [os.environ.get("CC", "gcc"), "-O2", "-std=c99", "-fPIC", "-shared", "-o", XTC_LIB, XTC_SRC, "-lm"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("gcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    return XTC_LIB


if __name__ == "__main__":
    print(build(force=True))
