import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs the live reference under /root/reference")
    # the unmodified reference calls numpy functions that are deprecated in numpy 2 (np.in1d in waterwire.py)
    config.addinivalue_line("filterwarnings", "ignore:.*in1d.*:DeprecationWarning")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    from oracle import ref_harness
    have_ref = ref_harness.reference_available()
    skip_ref = pytest.mark.skip(reason="live reference (/root/reference) not mounted")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
