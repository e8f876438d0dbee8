"""TEST INFRASTRUCTURE - ctypes front end of oracle/xtc_codec.c, the CPU restatement of the XTC frame codec.

The reference reads .xtc trajectories through MDAnalysis (cgraphs/mdhbond/basic.py:46-47), i.e. through MDAnalysis'
bundled xdrfile 1.1.4 (`xdrfile_decompress_coord_float`) followed by the float32 nm -> Angstrom conversion of
`MDAnalysis.coordinates.base` (`x *= 10.0`). None of that is in this image; `xtc_codec.c` restates the published
algorithm. PARITY UNPINNED (no golden .xtc file, no real reader to compare with) - see the header of xtc_codec.c.

Only tests/, `__graft_entry__.smoke()` and bench.py's CPU-baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        from oracle import build_oracle
        lib = C.CDLL(build_oracle.build())
        lib.xtc_oracle_frame_bound.restype = C.c_size_t
        lib.xtc_oracle_frame_bound.argtypes = [C.c_int]
        lib.xtc_oracle_encode.restype = C.c_size_t
        lib.xtc_oracle_encode.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        lib.xtc_oracle_frame_bytes.restype = C.c_size_t
        lib.xtc_oracle_frame_bytes.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
        lib.xtc_oracle_decode.restype = C.c_size_t
        lib.xtc_oracle_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _LIB = lib
    return _LIB


def encode(frames_nm, precision=1000.0, box_nm=None, first_step=0, dt=1.0):
    """XTC records of frames_nm float32 [F, N, 3] (nm): (uint8 array, int64 offsets [F + 1])."""
    lib = _lib()
    frames_nm = np.ascontiguousarray(frames_nm, dtype=np.float32)
    nf, na = frames_nm.shape[:2]
    box = np.zeros((nf, 9), np.float32) if box_nm is None else \
        np.ascontiguousarray(np.broadcast_to(np.asarray(box_nm, np.float32).reshape(-1, 9), (nf, 9)))
    bound = lib.xtc_oracle_frame_bound(na)
    buf = np.zeros(bound, dtype=np.uint8)
    parts, off = [], [0]
    for f in range(nf):
        n = lib.xtc_oracle_encode(frames_nm[f].ctypes.data, na, float(precision), first_step + f, float(dt * (first_step + f)),
                                  box[f].ctypes.data, buf.ctypes.data)
        if n == 0:
            raise ValueError("xtc oracle: frame %d cannot be encoded (coordinate out of range)" % f)
        parts.append(buf[:n].copy())
        off.append(off[-1] + n)
    return (np.concatenate(parts) if parts else np.zeros(0, np.uint8)), np.asarray(off, dtype=np.int64)


def index(buf):
    """(int64 offsets [F + 1], n_atoms) of the records in buf."""
    lib = _lib()
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    off, pos, na = [0], 0, C.c_int(0)
    while pos < buf.size:
        n = lib.xtc_oracle_frame_bytes(buf.ctypes.data + pos, buf.size - pos, C.byref(na))
        if n == 0:
            break
        pos += n
        off.append(pos)
    return np.asarray(off, dtype=np.int64), int(na.value)


def decode(buf, offsets, n_atoms, scale=10.0):
    """All frames: (float32 [F, N, 3] = fl32(fl32(int * fl32(1/precision)) * scale), float32 boxes [F, 3, 3] in nm)."""
    lib = _lib()
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    nf = len(offsets) - 1
    xyz = np.empty((nf, n_atoms, 3), dtype=np.float32)
    box = np.empty((nf, 9), dtype=np.float32)
    for f in range(nf):
        n = lib.xtc_oracle_decode(buf.ctypes.data + int(offsets[f]), int(offsets[f + 1] - offsets[f]), n_atoms, float(scale),
                                  xyz[f].ctypes.data, box[f].ctypes.data, None, None, None)
        if n != offsets[f + 1] - offsets[f]:
            raise ValueError("xtc oracle: frame %d is malformed" % f)
    return xyz, box.reshape(nf, 3, 3)
