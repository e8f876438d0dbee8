"""The C-ABI library builds, loads and exports every symbol include/hbgpu.h declares; without a GPU
the product path fails loudly instead of falling back to a CPU implementation."""
import ctypes
import os
import re

import pytest

from cgraphs_b200 import build as hb_build
from cgraphs_b200.mdhbond import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    hb_build.build()
    return _native.load_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hbgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 20
    bound = {n for n, _, _ in _native.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), "libhbgpu.so does not export %s" % n
        assert n in bound, "%s is declared in hbgpu.h but not bound in _native.py" % n
    assert bound <= set(names)


def test_struct_sizes_match_header(lib):
    # layout sanity for the by-pointer structs (pointers 8 bytes, natural alignment)
    assert ctypes.sizeof(_native.HbParams) == 40
    assert ctypes.sizeof(_native.HbTopology) == 8 + 3 * 8 + 8 + 5 * 8 + 8 + 2 * 8 + 8 + 4 * 8 + 8 + 8
    assert ctypes.sizeof(_native.HbTimers) == 16 + 16 * 8 + 16 * 8 + 5 * 8 + 8 + 32


def test_version_and_kernel_names(lib):
    assert b"sm_100a" in lib.hb_version()
    names = []
    k = 0
    while lib.hb_kernel_name(k):
        names.append(lib.hb_kernel_name(k).decode())
        k += 1
    assert "pairs" in names and "local_water" in names


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="this test documents the behaviour on a machine without a GPU")
def test_no_cpu_fallback(lib):
    h = ctypes.c_void_p()
    rc = lib.hb_create(ctypes.byref(h), 0)
    assert rc != 0 and not h.value
    assert lib.hb_last_error(None)
    from cgraphs_b200 import synth
    from cgraphs_b200.mdhbond import HbondAnalysis
    s = synth.make_system(900, 8, 1)
    hba = HbondAnalysis("protein", s.universe(1))      # topology tables are host work and succeed
    with pytest.raises(_native.HbError):
        hba.set_hbonds_in_selection()                  # the search itself needs the B200


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cgraphs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_header_is_plain_c(tmp_path):
    """include/hbgpu.h is the binding surface for any host language: it must compile as C99 on its own, and a C
    translation unit that references every entry point must link against libhbgpu.so."""
    import re
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hb_build.build()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "hbgpu.h")).read()
    names = sorted(set(re.findall(r"\b(hb_[a-z_0-9]+)\s*\(", header)))
    assert "hb_set_wires" in names and "hb_submit_frames" in names
    src = tmp_path / "use_abi.c"
    src.write_text('#include "hbgpu.h"\n#include <stdio.h>\nint main(void) {\n  const void* f[] = {%s};\n'
                   '  printf("%%d %%d\\n", (int)(sizeof f / sizeof f[0]), (int)sizeof(hb_timers));\n  return f[0] == 0;\n}\n'
                   % ", ".join("(const void*)%s" % n for n in names))
    exe = tmp_path / "use_abi"
    lib_dir = os.path.join(root, "cgraphs_b200", "csrc")
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                          "-L", lib_dir, "-lhbgpu", "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
