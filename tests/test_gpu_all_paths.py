"""Every kernel path once on small inputs (the script doubles as the compute-sanitizer driver)."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu


def test_all_kernel_paths_run():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "sanitize_small.py")
    spec = importlib.util.spec_from_file_location("sanitize_small", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()
