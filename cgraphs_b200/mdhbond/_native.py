"""ctypes binding of libhbgpu.so (include/hbgpu.h).  There is no CPU fallback: if the library or a
B200 is missing, the calls raise."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_LIB = None
LIB_PATH = os.environ.get("HBGPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc", "libhbgpu.so")

METHOD_IN_SELECTION = 0
METHOD_WATER_AROUND = 1
METHOD_ION_CONTACTS = 2
METHOD_WATER_PRESENCE = 3
PBC_NONE, PBC_ORTHO, PBC_TRICLINIC = 0, 1, 2


class HbTopology(C.Structure):
    _fields_ = [
        ("n_systems", C.c_int32),
        ("sys_atom_off", C.POINTER(C.c_int64)),
        ("sys_sel_off", C.POINTER(C.c_int32)),
        ("sys_wat_off", C.POINTER(C.c_int32)),
        ("n_sel", C.c_int32),
        ("sel_atom", C.POINTER(C.c_int32)),
        ("sel_don", C.POINTER(C.c_int32)),
        ("sel_acc", C.POINTER(C.c_int32)),
        ("sel_wat", C.POINTER(C.c_int32)),
        ("sel_bb", C.POINTER(C.c_uint8)),
        ("n_wat", C.c_int32),
        ("wat_atom", C.POINTER(C.c_int32)),
        ("wat_ent", C.POINTER(C.c_int32)),
        ("n_ent", C.c_int32),
        ("ent_group", C.POINTER(C.c_int32)),
        ("ent_resid", C.POINTER(C.c_int32)),
        ("ent_residue", C.POINTER(C.c_int32)),
        ("ent_h_off", C.POINTER(C.c_int32)),
        ("n_ent_h", C.c_int32),
        ("ent_h_atom", C.POINTER(C.c_int32)),
    ]


class HbParams(C.Structure):
    _fields_ = [
        ("distance", C.c_double),
        ("around_radius", C.c_double),
        ("cos_threshold", C.c_float),
        ("check_angle", C.c_int32),
        ("exclude_backbone_backbone", C.c_int32),
        ("method", C.c_int32),
        ("pbc_mode", C.c_int32),
        ("structure_batch", C.c_int32),
    ]


class HbTimers(C.Structure):
    _fields_ = [
        ("ms_total", C.c_double),
        ("ms_h2d", C.c_double),
        ("ms_kernel", C.c_double * 16),
        ("launches", C.c_int64 * 16),
        ("n_launches", C.c_int64),
        ("frames", C.c_int64),
        ("pairs_emitted", C.c_int64),
        ("table_capacity", C.c_int64),
        ("table_grows", C.c_int64),
        ("ms_run", C.c_double),
        ("fused_frames", C.c_int64),
        ("fused_fallbacks", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("finalize_retries", C.c_int64),
    ]


# every symbol include/hbgpu.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("hb_create", C.c_int, [C.POINTER(_P), C.c_int]),
    ("hb_destroy", C.c_int, [_P]),
    ("hb_last_error", C.c_char_p, [_P]),
    ("hb_version", C.c_char_p, []),
    ("hb_kernel_name", C.c_char_p, [C.c_int]),
    ("hb_set_topology", C.c_int, [_P, C.POINTER(HbTopology)]),
    ("hb_set_entry_groups", C.c_int, [_P, C.POINTER(C.c_int32), C.c_int32]),
    ("hb_set_params", C.c_int, [_P, C.POINTER(HbParams)]),
    ("hb_set_ions", C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_int32]),
    ("hb_set_wires", C.c_int, [_P, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("hb_set_wire_virtual_waters", C.c_int, [_P, _P, _P, _P, C.c_int64]),
    ("hb_set_wire_direct_nodes", C.c_int, [_P, _P, C.c_int32]),
    ("hb_set_option", C.c_int, [_P, C.c_char_p, C.c_int64]),
    ("hb_begin", C.c_int, [_P, C.c_int64]),
    ("hb_submit_frames", C.c_int, [_P, _P, C.c_int64, C.c_int64, _P]),
    ("hb_submit_frames_planar", C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P]),
    ("hb_submit_frames_device", C.c_int, [_P, _P, C.c_int64, C.c_int64, _P]),
    ("hb_submit_frames_xtc", C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_float, _P]),
    ("hb_xtc_decode_to_device", C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_float, _P]),
    ("hb_xtc_index", C.c_int, [_P, C.c_int64, _P, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("hb_xtc_decode_host", C.c_int, [_P, C.c_int64, C.c_int64, C.c_float, _P, _P, C.POINTER(C.c_int64)]),
    ("hb_xtc_encode", C.c_int, [_P, C.c_int64, C.c_int64, C.c_float, C.c_float, _P, C.c_int64, C.c_float, _P, C.c_int64, _P,
                                C.POINTER(C.c_int64), C.c_int32]),
    ("hb_finalize_async", C.c_int, [_P]),
    ("hb_pack_keys_device", C.c_int, [_P, _P, C.c_int32]),
    ("hb_finalize", C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("hb_get_keys", C.c_int, [_P, _P]),
    ("hb_get_bits", C.c_int, [_P, _P]),
    ("hb_get_keys_device", C.c_int, [_P, _P]),
    ("hb_get_bits_device", C.c_int, [_P, _P]),
    ("hb_merge_keys_device", C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P]),
    ("hb_comm_unique_id", C.c_int, [_P]),
    ("hb_comm_init", C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    ("hb_exchange_keys", C.c_int, [_P, C.c_int32, _P, _P]),
    ("hb_comm_destroy", C.c_int, [_P]),
    ("hb_get_counts", C.c_int, [_P, _P]),
    ("hb_joint_occupancy", C.c_int, [_P, _P, _P, C.POINTER(C.c_int64)]),
    ("hb_get_frame_column", C.c_int, [_P, C.c_int64, _P]),
    ("hb_get_timers", C.c_int, [_P, C.POINTER(HbTimers)]),
    ("hb_sync", C.c_int, [_P]),
    ("hb_host_alloc", C.c_int, [C.POINTER(_P), C.c_uint64]),
    ("hb_host_free", C.c_int, [_P]),
]


class HbError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libhbgpu.so and declare the prototypes.  Raises if the library was not built."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = os.path.abspath(path or LIB_PATH)
    if not os.path.exists(p):
        raise HbError("libhbgpu.so not found at %s - build it with `python -m cgraphs_b200.build` "
                      "(there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _LIB = lib
    return lib


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context:
    """One hb_ctx. Keeps the numpy arrays of the topology alive for the duration of the upload only."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _P()
        rc = self.lib.hb_create(C.byref(h), int(device))
        if rc != 0:
            raise HbError("hb_create failed (%d): %s" % (rc, (self.lib.hb_last_error(None) or b"").decode()))
        self.h = h
        self.device = int(device)

    def _ck(self, rc, what):
        if rc != 0:
            raise HbError("%s failed (%d): %s" % (what, rc, (self.lib.hb_last_error(self.h) or b"").decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.hb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        self._ck(self.lib.hb_set_option(self.h, name.encode(), int(value)), "hb_set_option")

    def set_topology(self, t):
        """t: dict of numpy arrays named like the hb_topology fields."""
        self._uploaded = None          # callers that cache packed tables set this marker afterwards
        keep = {}
        st = HbTopology()
        st.n_systems = int(t["n_systems"])
        keep["sys_atom_off"] = np.ascontiguousarray(t["sys_atom_off"], dtype=np.int64)
        st.sys_atom_off = _ptr(keep["sys_atom_off"], C.c_int64)
        for k in ("sys_sel_off", "sys_wat_off", "sel_atom", "sel_don", "sel_acc", "sel_wat", "wat_atom", "wat_ent",
                  "ent_group", "ent_resid", "ent_residue", "ent_h_off", "ent_h_atom"):
            keep[k] = _i32(t[k])
            setattr(st, k, _ptr(keep[k], C.c_int32))
        keep["sel_bb"] = np.ascontiguousarray(t["sel_bb"], dtype=np.uint8)
        st.sel_bb = _ptr(keep["sel_bb"], C.c_uint8)
        st.n_sel = len(keep["sel_atom"])
        st.n_wat = len(keep["wat_atom"])
        st.n_ent = len(keep["ent_group"])
        st.n_ent_h = len(keep["ent_h_atom"])
        assert len(keep["ent_h_off"]) == st.n_ent + 1
        self._ck(self.lib.hb_set_topology(self.h, C.byref(st)), "hb_set_topology")

    def set_ions(self, ion_atom, ion_group, include_water=True):
        a, g = _i32(ion_atom), _i32(ion_group)
        self._ck(self.lib.hb_set_ions(self.h, _ptr(a, C.c_int32), _ptr(g, C.c_int32), len(a), int(bool(include_water))),
                 "hb_set_ions")

    def set_wires(self, ent_node, n_res, max_water, allow_direct_bonds=True):
        """Water-wire mode (hb_set_wires); ent_node=None switches it off."""
        if ent_node is None:
            self._ck(self.lib.hb_set_wires(self.h, None, 0, 0, 0, 0), "hb_set_wires")
            return
        g = _i32(ent_node)
        self._ck(self.lib.hb_set_wires(self.h, _ptr(g, C.c_int32), len(g), int(n_res), int(max_water),
                                       int(bool(allow_direct_bonds))), "hb_set_wires")

    def set_wire_direct_nodes(self, direct_node=None):
        if direct_node is None:
            self._ck(self.lib.hb_set_wire_direct_nodes(self.h, None, 0), "hb_set_wire_direct_nodes")
            return
        d = np.ascontiguousarray(direct_node, dtype=np.int32)
        self._ck(self.lib.hb_set_wire_direct_nodes(self.h, d.ctypes.data, len(d)), "hb_set_wire_direct_nodes")

    def set_wire_virtual_waters(self, frame_off=None, true_wat=None, label_wat=None):
        """Per-frame (position water, name water) lists of the convex-hull wire option; None switches it off."""
        if frame_off is None:
            self._ck(self.lib.hb_set_wire_virtual_waters(self.h, None, None, None, 0), "hb_set_wire_virtual_waters")
            return
        off = np.ascontiguousarray(frame_off, dtype=np.int64)
        t = np.ascontiguousarray(true_wat, dtype=np.int32)
        l = np.ascontiguousarray(label_wat, dtype=np.int32)
        self._ck(self.lib.hb_set_wire_virtual_waters(self.h, off.ctypes.data, t.ctypes.data if len(t) else None,
                                                     l.ctypes.data if len(l) else None, len(off) - 1),
                 "hb_set_wire_virtual_waters")

    def set_entry_groups(self, groups):
        g = _i32(groups)
        self._ck(self.lib.hb_set_entry_groups(self.h, _ptr(g, C.c_int32), len(g)), "hb_set_entry_groups")

    def set_params(self, distance, around_radius, cos_threshold, check_angle, exclude_backbone_backbone, method,
                   pbc_mode=0, structure_batch=False):
        p = HbParams(float(distance), float(around_radius), float(cos_threshold), int(bool(check_angle)),
                     int(bool(exclude_backbone_backbone)), int(method), int(pbc_mode), int(bool(structure_batch)))
        self._ck(self.lib.hb_set_params(self.h, C.byref(p)), "hb_set_params")

    def begin(self, n_frames_total):
        self._run_id = getattr(self, "_run_id", 0) + 1     # device-side results of earlier runs are gone
        self._ck(self.lib.hb_begin(self.h, int(n_frames_total)), "hb_begin")

    def submit(self, xyz, n_atoms=None, box=None):
        """xyz: float32 host array [F, N, 3]; box: None or float32 [F, 3, 3] lattice rows (pbc modes)."""
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        if a.ndim == 3:
            nf, na = a.shape[0], a.shape[1]
        else:
            raise ValueError("submit expects [F, N, 3]; use submit_raw for ragged batches")
        b = None
        if box is not None:
            b = np.ascontiguousarray(box, dtype=np.float32).reshape(nf, 9)
        self._ck(self.lib.hb_submit_frames(self.h, a.ctypes.data, na, nf, b.ctypes.data if b is not None else None),
                 "hb_submit_frames")

    def submit_raw(self, ptr, n_atoms, n_frames, box=None):
        b = None
        if box is not None:
            b = np.ascontiguousarray(box, dtype=np.float32).reshape(int(n_frames), 9)
        self._ck(self.lib.hb_submit_frames(self.h, int(ptr), int(n_atoms), int(n_frames),
                                           b.ctypes.data if b is not None else None), "hb_submit_frames")

    def submit_planar(self, raw, frame_stride, x_off, y_off, z_off, n_atoms, n_frames, box=None):
        """raw: contiguous uint8 array (or address) of n_frames frame records with X / Y / Z planes (DCD layout)."""
        b = None
        if box is not None:
            b = np.ascontiguousarray(box, dtype=np.float32).reshape(int(n_frames), 9)
        ptr = raw if isinstance(raw, int) else raw.ctypes.data
        self._ck(self.lib.hb_submit_frames_planar(self.h, ptr, int(frame_stride), int(x_off), int(y_off), int(z_off),
                                                  int(n_atoms), int(n_frames), b.ctypes.data if b is not None else None),
                 "hb_submit_frames_planar")

    def submit_xtc(self, raw, frame_off, n_atoms, n_frames, scale=10.0, box=None):
        """raw: contiguous uint8 array (or address) holding the XTC records; frame_off: int64 [n_frames + 1] offsets."""
        b = None
        if box is not None:
            b = np.ascontiguousarray(box, dtype=np.float32).reshape(int(n_frames), 9)
        off = np.ascontiguousarray(frame_off, dtype=np.int64)
        if len(off) != int(n_frames) + 1:
            raise ValueError("frame_off must hold n_frames + 1 offsets")
        ptr = raw if isinstance(raw, int) else raw.ctypes.data
        self._ck(self.lib.hb_submit_frames_xtc(self.h, ptr, off.ctypes.data, int(n_atoms), int(n_frames), float(scale),
                                               b.ctypes.data if b is not None else None), "hb_submit_frames_xtc")

    def xtc_decode_to_device(self, raw, frame_off, n_atoms, n_frames, dev_ptr, scale=10.0):
        off = np.ascontiguousarray(frame_off, dtype=np.int64)
        ptr = raw if isinstance(raw, int) else raw.ctypes.data
        self._ck(self.lib.hb_xtc_decode_to_device(self.h, ptr, off.ctypes.data, int(n_atoms), int(n_frames), float(scale),
                                                  int(dev_ptr)), "hb_xtc_decode_to_device")

    def submit_device(self, dev_ptr, n_atoms, n_frames, box_dev_ptr=None):
        self._ck(self.lib.hb_submit_frames_device(self.h, int(dev_ptr), int(n_atoms), int(n_frames),
                                                  int(box_dev_ptr) if box_dev_ptr else None),
                 "hb_submit_frames_device")

    def sync(self):
        self._ck(self.lib.hb_sync(self.h), "hb_sync")

    def finalize_async(self):
        self._ck(self.lib.hb_finalize_async(self.h), "hb_finalize_async")

    def pack_keys_device(self, block_ptr, cap):
        self._ck(self.lib.hb_pack_keys_device(self.h, int(block_ptr), int(cap)), "hb_pack_keys_device")

    def finalize(self):
        nb, wpr = C.c_int64(), C.c_int64()
        self._ck(self.lib.hb_finalize(self.h, C.byref(nb), C.byref(wpr)), "hb_finalize")
        return nb.value, wpr.value

    def results(self):
        """(keys uint64 [B] ascending, words uint32 [B, wpr])."""
        nb, wpr = self.finalize()
        keys = np.empty(nb, dtype=np.uint64)
        words = np.empty((nb, wpr), dtype=np.uint32)
        if nb:
            self._ck(self.lib.hb_get_keys(self.h, keys.ctypes.data), "hb_get_keys")
            self._ck(self.lib.hb_get_bits(self.h, words.ctypes.data), "hb_get_bits")
        return keys, words

    def keys_to_device(self, dev_ptr, sync=True):
        """Queue the copy on the context's compute stream; sync=False when the caller works on that stream."""
        self._ck(self.lib.hb_get_keys_device(self.h, int(dev_ptr)), "hb_get_keys_device")
        if sync:
            self.sync()

    def bits_to_device(self, dev_ptr, sync=True):
        self._ck(self.lib.hb_get_bits_device(self.h, int(dev_ptr)), "hb_get_bits_device")
        if sync:
            self.sync()

    def merge_keys_device(self, gathered_ptr, world, cap, out_ptr, info_ptr):
        self._ck(self.lib.hb_merge_keys_device(self.h, int(gathered_ptr), int(world), int(cap), int(out_ptr), int(info_ptr)),
                 "hb_merge_keys_device")

    def comm_init(self, unique_id, world, rank):
        """hb_comm_init: join the NCCL communicator of the key exchange (unique_id: the 128 bytes of comm_unique_id())."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.hb_comm_init(self.h, buf, int(world), int(rank)), "hb_comm_init")

    def exchange_keys(self, cap, out_ptr, info_ptr):
        self._ck(self.lib.hb_exchange_keys(self.h, int(cap), int(out_ptr), int(info_ptr)), "hb_exchange_keys")

    def counts(self, n_bonds):
        c = np.zeros(n_bonds, dtype=np.int64)
        if n_bonds:
            self._ck(self.lib.hb_get_counts(self.h, c.ctypes.data), "hb_get_counts")
        return c

    def joint_occupancy(self, wpr, select=None):
        """(uint32 words [wpr], number of frames in which every selected bond exists)."""
        words = np.zeros(wpr, dtype=np.uint32)
        cnt = C.c_int64()
        sel = None
        if select is not None:
            sel = np.ascontiguousarray(select, dtype=np.uint8)
        self._ck(self.lib.hb_joint_occupancy(self.h, sel.ctypes.data if sel is not None else None, words.ctypes.data,
                                             C.byref(cnt)), "hb_joint_occupancy")
        return words, cnt.value

    def frame_column(self, frame, n_bonds):
        present = np.zeros(n_bonds, dtype=np.uint8)
        if n_bonds:
            self._ck(self.lib.hb_get_frame_column(self.h, int(frame), present.ctypes.data), "hb_get_frame_column")
        return present.astype(bool)

    def timers(self):
        t = HbTimers()
        self._ck(self.lib.hb_get_timers(self.h, C.byref(t)), "hb_get_timers")
        names = []
        k = 0
        while True:
            n = self.lib.hb_kernel_name(k)
            if not n:
                break
            names.append(n.decode())
            k += 1
        return dict(ms_total=t.ms_total, ms_h2d=t.ms_h2d, ms_run=t.ms_run, n_launches=t.n_launches, frames=t.frames,
                    pairs_emitted=t.pairs_emitted, table_capacity=t.table_capacity, table_grows=t.table_grows,
                    fused_frames=t.fused_frames, fused_fallbacks=t.fused_fallbacks, h2d_bytes=t.h2d_bytes,
                    finalize_retries=t.finalize_retries,
                    ms_kernel={n: t.ms_kernel[i] for i, n in enumerate(names)},
                    launches={n: t.launches[i] for i, n in enumerate(names)})


class PinnedBuffer:
    """numpy array (float32 unless `dtype` says otherwise) over cudaHostAlloc memory."""

    def __init__(self, shape, dtype=np.float32):
        lib = load_library()
        self.lib = lib
        dt = np.dtype(dtype)
        n = int(np.prod(shape))
        p = _P()
        rc = lib.hb_host_alloc(C.byref(p), max(n * dt.itemsize, 16))
        if rc != 0:
            raise HbError("hb_host_alloc(%d bytes) failed" % (n * dt.itemsize))
        self.ptr = p.value
        buf = (C.c_uint8 * (n * dt.itemsize)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dt).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.hb_host_free(self.ptr)
            self.ptr = None


# ------------------------------------------------------------------------------------------------
# XTC host helpers (no context, no GPU): index, single-frame decode, writer
# ------------------------------------------------------------------------------------------------
def xtc_index(buf):
    """(int64 offsets [n_frames + 1], n_atoms) of the XTC frame records in the uint8 array `buf`."""
    lib = load_library()
    buf = np.asarray(buf)
    n, na = C.c_int64(), C.c_int64()
    ptr = buf.ctypes.data if buf.size else None
    if lib.hb_xtc_index(ptr, buf.size, None, 0, C.byref(n), C.byref(na)) != 0:
        raise HbError("hb_xtc_index: not an XTC file (or the atom count changes between frames)")
    off = np.zeros(n.value + 1, dtype=np.int64)
    if n.value:
        lib.hb_xtc_index(ptr, buf.size, off.ctypes.data, n.value, C.byref(n), C.byref(na))
    return off, int(na.value)


def xtc_decode_host(buf, offset, n_atoms, scale=10.0):
    """(float32 [n_atoms, 3], float32 box rows [3, 3] in nm) of the record at buf[offset:], decoded on the CPU."""
    lib = load_library()
    buf = np.asarray(buf)
    xyz = np.empty((int(n_atoms), 3), dtype=np.float32)
    box = np.empty(9, dtype=np.float32)
    if lib.hb_xtc_decode_host(buf.ctypes.data + int(offset), buf.size - int(offset), int(n_atoms), float(scale),
                              xyz.ctypes.data, box.ctypes.data, None) != 0:
        raise HbError("hb_xtc_decode_host: malformed XTC record")
    return xyz, box.reshape(3, 3)


def xtc_encode(xyz, in_scale=0.1, precision=1000.0, box_nm=None, first_step=0, dt=1.0, n_threads=0):
    """XTC records of the frames xyz [F, N, 3] (times in_scale = nm): (uint8 array, int64 offsets [F + 1])."""
    lib = load_library()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    nf, na = xyz.shape[0], xyz.shape[1]
    b = None
    if box_nm is not None:
        b = np.ascontiguousarray(np.broadcast_to(np.asarray(box_nm, dtype=np.float32).reshape(-1, 9), (nf, 9)))
    if n_threads <= 0:
        n_threads = min(64, os.cpu_count() or 1)
    cap = nf * (96 + 15 * na + 64)
    out = np.empty(cap, dtype=np.uint8)
    off = np.zeros(nf + 1, dtype=np.int64)
    written = C.c_int64()
    rc = lib.hb_xtc_encode(xyz.ctypes.data, na, nf, float(in_scale), float(precision), b.ctypes.data if b is not None else None,
                           int(first_step), float(dt), out.ctypes.data, cap, off.ctypes.data, C.byref(written), int(n_threads))
    if rc != 0:
        raise HbError("hb_xtc_encode failed (%d): coordinate out of range?" % rc)
    return out[:written.value].copy() if written.value < cap // 2 else out[:written.value], off


def xtc_decode_device(buf, offsets, n_atoms, scale=10.0, device=0):
    """Decode XTC records on the GPU and return the frames as a numpy array (tests, spot checks); needs torch."""
    import torch
    nf = len(offsets) - 1
    out = torch.empty((nf, int(n_atoms), 3), dtype=torch.float32, device="cuda:%d" % device)
    ctx = Context(device)
    try:
        ctx.xtc_decode_to_device(np.ascontiguousarray(buf, dtype=np.uint8), offsets, n_atoms, nf, out.data_ptr(), scale)
    finally:
        ctx.close()
    return out.cpu().numpy()


def comm_unique_id():
    """128-byte NCCL id (hb_comm_unique_id) that rank 0 creates and every rank passes to Context.comm_init."""
    lib = load_library()
    buf = (C.c_char * 128)()
    if lib.hb_comm_unique_id(buf) != 0:
        raise HbError("hb_comm_unique_id failed: " + (lib.hb_last_error(None) or b"").decode())
    return bytes(buf)
