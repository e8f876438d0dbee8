"""TEST INFRASTRUCTURE - "builds" the reference for the GPU box: byte-compiles the reference's own Python sources,
from where they lie under /root/reference, into oracle/_ref/ (git-ignored, travels with the snapshot).

    python oracle/make_ref.py            (also called by __graft_entry__.build() when /root/reference exists)

The reference (evabertalan/cgraphs) is pure Python, so its analogue of compiling a C reference from its own source
files is `py_compile`: oracle/_ref/cgraphs/mdhbond/*.pyc and oracle/_ref/cgraphs/{conservedgraph,helperfunctions,
proteingraphanalyser}.pyc are compiler OUTPUT; no reference source file is copied into the repository. The bytecode is
imported unmodified through oracle/ref_harness.py (MDAnalysis / matplotlib stubbed, as in the CPU tests) by
`bench.py --impl reference` (cpu_baseline.kind = "reference") and by nothing in the product.
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CGRAPHS_REF_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["cgraphs/mdhbond/__init__.py", "cgraphs/mdhbond/basic.py", "cgraphs/mdhbond/hbond.py",
         "cgraphs/mdhbond/helpfunctions.py", "cgraphs/mdhbond/hydration.py", "cgraphs/mdhbond/network.py",
         "cgraphs/mdhbond/waterwire.py", "cgraphs/conservedgraph.py", "cgraphs/helperfunctions.py",
         "cgraphs/proteingraphanalyser.py"]


def available():
    return os.path.isfile(os.path.join(REF, "cgraphs", "mdhbond", "hbond.py"))


def build():
    if not available():
        return None
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, rel + "c.bin")       # (.pyc files are filtered out of the snapshot that goes to the GPU box)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, doraise=True, optimize=0)
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as f:
        f.write("byte-compiled from %s by oracle/make_ref.py with Python %s; compiler output, not source\n" % (REF, sys.version.split()[0]))
    return OUT


if __name__ == "__main__":
    print(build())
