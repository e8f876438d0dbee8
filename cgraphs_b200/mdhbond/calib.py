"""Host calibration of the angle criterion.

The reference decides ``rad2deg(arccos(cos_arg)) <= cut_angle`` in float32 with numpy
(``cgraphs/mdhbond/helpfunctions.py:102-104,155``).  numpy's float32 ``arccos`` is a SIMD kernel that is
not correctly rounded, so the GPU cannot reproduce it from first principles; but the composite is
monotone non-increasing in ``cos_arg``, so the test is equivalent to ``c_star <= cos_arg <= 1`` where
``c_star`` is the smallest float32 that still passes *with this host's numpy*.  It is found here by
bisection over the float32 bit patterns and the monotonicity is verified in a window around it.
"""
from __future__ import annotations

import functools
import numpy as np


def _passes(c, cut32):
    with np.errstate(all="ignore"):
        return np.rad2deg(np.arccos(np.asarray(c, dtype=np.float32))) <= cut32


def _ordered(bits):
    """int32 bit pattern -> monotone integer key of the float ordering (and back, it is an involution
    on the sign-magnitude part)."""
    bits = np.asarray(bits, dtype=np.int64)
    return np.where(bits < 0, -(bits & 0x7FFFFFFF), bits)


def _from_key(key):
    key = np.asarray(key, dtype=np.int64)
    bits = np.where(key < 0, (-key) | 0x80000000, key).astype(np.uint32)
    return bits.view(np.float32)


@functools.lru_cache(maxsize=64)
def cos_threshold(cut_angle, window=1 << 15):
    """Smallest float32 ``c`` in [-1, 1] with ``rad2deg(arccos(c)) <= float32(cut_angle)``.

    Returns 2.0 when no value passes (negative cut-off).  ``cut_angle`` is compared as the reference
    does: a Python float against a float32 array is a float32 comparison under numpy >= 2.
    """
    cut32 = np.float32(cut_angle)
    if not bool(_passes(np.float32(1.0), cut32)):
        return float(2.0)
    lo_key = int(_ordered(np.float32(-1.0).view(np.int32)))
    hi_key = int(_ordered(np.float32(1.0).view(np.int32)))
    if bool(_passes(np.float32(-1.0), cut32)):
        return float(-1.0)
    # invariant: key lo fails, key hi passes
    lo, hi = lo_key, hi_key
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if bool(_passes(_from_key(mid), cut32)):
            hi = mid
        else:
            lo = mid
    # verify monotonicity around the threshold: everything below fails, everything at/above passes
    keys = np.arange(max(lo_key, hi - window), min(hi_key, hi + window) + 1, dtype=np.int64)
    ok = _passes(_from_key(keys), cut32)
    expect = keys >= hi
    if not np.array_equal(ok, expect):
        raise RuntimeError("angle criterion is not monotone near the cut-off on this host's numpy; "
                           "cannot derive a cosine threshold for cut_angle=%r" % (cut_angle,))
    return float(_from_key(hi))
