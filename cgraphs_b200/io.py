"""Structure / trajectory files without MDAnalysis: PDB topologies and CHARMM/NAMD DCD trajectories.

The reference gets its coordinates from ``MDAnalysis.Universe(structure, trajectories)``
(``cgraphs/mdhbond/basic.py:46-47``); MDAnalysis is not installable in this image, and its DCD reader decodes
every frame on the CPU (de-interleaving the X / Y / Z records). Here the frame records of a DCD file are handed
to ``libhbgpu`` as they lie in the file (``hb_submit_frames_planar``): they cross PCIe untouched and are
interleaved on the device.

* :func:`read_pdb` / :func:`write_pdb`  - topology columns (names, resnames, resids, segids), one coordinate set,
  ``CRYST1`` cell
* :func:`read_psf` / :func:`write_psf`  - CHARMM / NAMD / X-PLOR PSF topologies (what cgraphs' trajectory workflow
  passes as ``structure``, ``proteingraphanalyser.py:472-476``); coordinates then come from the trajectory
* :class:`DCDFrames` / :func:`write_dcd` - frame source over a DCD file (memory mapped), with the planar view
  ``planar_block`` the native path uses, and the unit cell of every frame
* :class:`XTCFrames` / :func:`write_xtc` - frame source over a GROMACS XTC file: the compressed frame records are
  handed to ``hb_submit_frames_xtc`` as they lie in the file and decoded on the device; coordinates come out in
  Angstrom with the float32 arithmetic of MDAnalysis' reader (xdrfile decode, then ``x *= 10``)
* :func:`universe_from_files` - ``Universe`` (cgraphs_b200.universe) from a PDB / PSF and optional DCD / XTC files
"""
from __future__ import annotations

import struct

import numpy as np

from .universe import Topology, Universe


# ------------------------------------------------------------------------------------------------
# PDB
# ------------------------------------------------------------------------------------------------
def write_pdb(path, topology, xyz, dimensions=None):
    """Fixed-column PDB with 4-character residue names (CHARMM style: TIP3) and the segid in columns 73-76."""
    with open(path, "w") as f:
        if dimensions is not None:
            a, b, c, al, be, ga = [float(x) for x in dimensions]
            f.write("CRYST1%9.3f%9.3f%9.3f%7.2f%7.2f%7.2f P 1           1\n" % (a, b, c, al, be, ga))
        for i in range(topology.n_atoms):
            name = str(topology.names[i])
            name_field = (" %-3s" % name) if len(name) < 4 else name[:4]
            f.write("ATOM  %5d %4s %-4s%1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f      %-4s\n" % (
                (i + 1) % 100000, name_field, str(topology.resnames[i])[:4], " ", int(topology.resids[i]) % 10000,
                xyz[i, 0], xyz[i, 1], xyz[i, 2], 1.0, 0.0, str(topology.segids[i])[:4]))
        f.write("END\n")


def read_pdb(path):
    """Returns (Topology, float32 [N, 3], dimensions or None). ATOM / HETATM records, first model only.

    Residue numbers above 9999 wrap in the file; a new residue starts whenever (segid, resid, resname) changes,
    and wrapped numbers are unwrapped per segment so that residues stay distinct."""
    names, resnames, resids, segids, xyz, types = [], [], [], [], [], []
    dims = None
    with open(path) as f:
        for line in f:
            rec = line[:6]
            if rec == "CRYST1":
                dims = [float(line[6:15]), float(line[15:24]), float(line[24:33]), float(line[33:40]), float(line[40:47]),
                        float(line[47:54])]
            elif rec in ("ATOM  ", "HETATM"):
                names.append(line[12:16].strip())
                resnames.append(line[17:21].strip())
                resids.append(int(line[22:26]))
                seg = line[72:76].strip() if len(line) >= 76 else ""
                segids.append(seg if seg else (line[21].strip() or "SYSTEM"))
                xyz.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
                # atom type like MDAnalysis' PDB parser: the element columns 77-78 when present, else guessed from the
                # name (leading digits dropped: '1HB' is a hydrogen) - the reference's donor rule looks at it
                # (helpfunctions.py:57-66)
                el = line[76:78].strip().upper() if len(line) >= 78 else ""
                if not el:
                    stripped = names[-1].lstrip("0123456789")
                    el = stripped[:1].upper()
                types.append(el)
            elif rec.startswith("ENDMDL"):
                break
    resids = np.asarray(resids, dtype=np.int64)
    # unwrap residue numbers that restarted after 9999 inside one segment
    out = resids.copy()
    offset, prev_seg, prev_raw = 0, None, None
    for i, (seg, r) in enumerate(zip(segids, resids)):
        if seg != prev_seg:
            offset = 0
        elif r < prev_raw and prev_raw - r > 5000:
            offset += 10000
        out[i] = r + offset
        prev_seg, prev_raw = seg, r
    return Topology(names, resnames, out, segids, types=types), np.asarray(xyz, dtype=np.float32), dims


# ------------------------------------------------------------------------------------------------
# PSF (CHARMM / NAMD / X-PLOR; standard and EXT column widths - the atom records are read field by field)
# ------------------------------------------------------------------------------------------------
def read_psf(path):
    """Topology of the !NATOM section: segid, resid, resname, atom name and the PSF atom type (MDAnalysis' PSF parser
    exposes that type string as ``atom.type``; the reference's hydrogen rule reads it, helpfunctions.py:57-66)."""
    names, resnames, resids, segids, types = [], [], [], [], []
    with open(path) as f:
        first = f.readline()
        if not first.startswith("PSF"):
            raise ValueError("%s: not a PSF file" % path)
        n_atoms = None
        for line in f:
            if "!NATOM" in line:
                n_atoms = int(line.split()[0])
                break
        if n_atoms is None:
            raise ValueError("%s: no !NATOM section" % path)
        for _ in range(n_atoms):
            parts = f.readline().split()
            if len(parts) < 6:
                raise ValueError("%s: truncated atom record" % path)
            segids.append(parts[1])
            digits = parts[2].lstrip("-")
            k = 0
            while k < len(digits) and digits[k].isdigit():
                k += 1
            resids.append(int(parts[2][:len(parts[2]) - len(digits) + k]))      # '12A' (insertion code) -> 12
            resnames.append(parts[3])
            names.append(parts[4])
            types.append(parts[5])
    return Topology(names, resnames, np.asarray(resids, dtype=np.int64), segids, types=types)


def write_psf(path, topology, types=None):
    """Minimal EXT PSF with the !NATOM section only (tests, and conversion of generated systems)."""
    types = topology.types if types is None else types
    with open(path, "w") as f:
        f.write("PSF EXT\n\n%10d !NTITLE\n REMARKS written by cgraphs_b200.io\n\n%10d !NATOM\n" % (1, topology.n_atoms))
        for i in range(topology.n_atoms):
            f.write("%10d %-8s %-8d %-8s %-8s %-6s %10.6f %13.4f %11d\n" % (
                i + 1, str(topology.segids[i]), int(topology.resids[i]), str(topology.resnames[i]), str(topology.names[i]),
                str(types[i]) or "X", 0.0, 1.0, 0))
        f.write("\n%10d !NBOND: bonds\n\n" % 0)


# ------------------------------------------------------------------------------------------------
# DCD (CHARMM / NAMD, little endian, 32-bit record markers)
# ------------------------------------------------------------------------------------------------
def write_dcd(path, frames, dimensions=None):
    """frames: float32 array [F, N, 3] (or an iterable of [N, 3]); dimensions: [a, b, c, alpha, beta, gamma] or None."""
    frames = [np.asarray(fr, dtype=np.float32) for fr in frames]
    n_frames, n_atoms = len(frames), frames[0].shape[0]
    with open(path, "wb") as f:
        icntrl = [0] * 20
        icntrl[0], icntrl[1], icntrl[2] = n_frames, 0, 1
        icntrl[10] = 1 if dimensions is not None else 0
        icntrl[19] = 24                                   # CHARMM version: unit cell block holds six doubles
        hdr = b"CORD" + struct.pack("<9i", *icntrl[:9]) + struct.pack("<f", 1.0) + struct.pack("<10i", *icntrl[10:])
        f.write(struct.pack("<i", 84) + hdr + struct.pack("<i", 84))
        title = b"written by cgraphs_b200.io".ljust(80)
        f.write(struct.pack("<i", 84) + struct.pack("<i", 1) + title + struct.pack("<i", 84))
        f.write(struct.pack("<3i", 4, n_atoms, 4))
        for fr in frames:
            if dimensions is not None:
                a, b, c, al, be, ga = [float(x) for x in dimensions]
                cell = struct.pack("<6d", a, np.cos(np.deg2rad(ga)), b, np.cos(np.deg2rad(be)), np.cos(np.deg2rad(al)), c)
                f.write(struct.pack("<i", 48) + cell + struct.pack("<i", 48))
            for k in range(3):
                f.write(struct.pack("<i", 4 * n_atoms))
                f.write(np.ascontiguousarray(fr[:, k], dtype="<f4").tobytes())
                f.write(struct.pack("<i", 4 * n_atoms))


class DCDFrames:
    """Frame source over a DCD file: ``len``, ``frame(i)``, ``block(a, b)`` like the array sources of
    cgraphs_b200.universe, plus ``planar_block(a, b)`` (the raw frame records and where the X / Y / Z planes lie in
    them) for ``hb_submit_frames_planar`` and ``dimensions_of(i)``."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            head = f.read(92)
            if len(head) < 92 or struct.unpack("<i", head[:4])[0] != 84 or head[4:8] != b"CORD":
                raise ValueError("%s: not a little-endian CHARMM/NAMD DCD file" % path)
            icntrl = struct.unpack("<9i", head[8:44]) + (0,) + struct.unpack("<10i", head[48:88])
            self.n_frames_header = icntrl[0]
            if icntrl[8] != 0:
                raise ValueError("%s: DCD files with fixed atoms are not supported" % path)
            self.has_cell = icntrl[10] != 0
            if icntrl[11] != 0:
                raise ValueError("%s: 4-dimensional DCD files are not supported" % path)
            self.charmm_version = icntrl[19]
            n = struct.unpack("<i", f.read(4))[0]
            f.seek(n, 1)
            if struct.unpack("<i", f.read(4))[0] != n:
                raise ValueError("%s: corrupt title block" % path)
            a, self.n_atoms, b = struct.unpack("<3i", f.read(12))
            if a != 4 or b != 4:
                raise ValueError("%s: corrupt atom-count block" % path)
            self.data_offset = f.tell()
        plane = 4 * self.n_atoms + 8
        self.cell_bytes = 56 if self.has_cell else 0
        self.frame_stride = self.cell_bytes + 3 * plane
        self.x_off = self.cell_bytes + 4
        self.y_off = self.x_off + plane
        self.z_off = self.y_off + plane
        self._map = np.memmap(path, dtype=np.uint8, mode="r", offset=self.data_offset)
        self.n_frames = len(self._map) // self.frame_stride
        if self.n_frames_header and self.n_frames_header < self.n_frames:
            self.n_frames = self.n_frames_header

    def __len__(self):
        return self.n_frames

    def planar_block(self, start, stop):
        """(uint8 view of the frame records [start, stop), frame_stride, x_off, y_off, z_off)."""
        raw = self._map[start * self.frame_stride:stop * self.frame_stride]
        return raw, self.frame_stride, self.x_off, self.y_off, self.z_off

    def planar_count(self, start, stop):
        return stop - start

    def block(self, start, stop):
        raw = np.asarray(self._map[start * self.frame_stride:stop * self.frame_stride]).reshape(stop - start, self.frame_stride)
        out = np.empty((stop - start, self.n_atoms, 3), dtype=np.float32)
        for k, off in enumerate((self.x_off, self.y_off, self.z_off)):
            out[:, :, k] = raw[:, off:off + 4 * self.n_atoms].copy().view("<f4")
        return out

    def frame(self, i):
        return self.block(int(i), int(i) + 1)[0]

    def dimensions_of(self, i):
        """[a, b, c, alpha, beta, gamma] of frame i, or None."""
        if not self.has_cell:
            return None
        o = i * self.frame_stride + 4
        a, g, b, be, al, c = np.asarray(self._map[o:o + 48]).view("<f8")
        ang = []
        for v in (al, be, g):      # CHARMM >= 25 stores cosines, older files and NAMD may store degrees
            ang.append(float(np.rad2deg(np.arccos(v))) if -1.0 <= v <= 1.0 else float(v))
        return [float(a), float(b), float(c), ang[0], ang[1], ang[2]]


class ConcatDCDFrames:
    """Several DCD files of one system read as one trajectory (MDAnalysis' ChainReader for the list of trajectories
    cgraphs passes, ``proteingraphanalyser.py:472-476``). ``planar_block`` hands out frame records of ONE file, so a
    block that would cross a file boundary ends there (callers advance by the number of frames they got)."""

    def __init__(self, paths):
        self.parts = [DCDFrames(p) for p in paths]
        if not self.parts:
            raise ValueError("no trajectory files")
        self.n_atoms = self.parts[0].n_atoms
        for p, part in zip(paths, self.parts):
            if part.n_atoms != self.n_atoms:
                raise ValueError("%s has %d atoms, %s has %d" % (p, part.n_atoms, paths[0], self.n_atoms))
        self.offsets = np.concatenate([[0], np.cumsum([len(p) for p in self.parts])]).astype(np.int64)
        self.n_frames = int(self.offsets[-1])
        self.has_cell = all(p.has_cell for p in self.parts)

    def __len__(self):
        return self.n_frames

    def _locate(self, i):
        k = int(np.searchsorted(self.offsets, i, side="right") - 1)
        return k, int(i - self.offsets[k])

    def planar_block(self, start, stop):
        k, a = self._locate(start)
        b = min(int(stop - self.offsets[k]), len(self.parts[k]))
        return self.parts[k].planar_block(a, b)

    def planar_count(self, start, stop):
        """Frames `planar_block(start, stop)` returns (fewer than stop - start at a file boundary)."""
        k, a = self._locate(start)
        return min(int(stop - self.offsets[k]), len(self.parts[k])) - a

    def block(self, start, stop):
        out = np.empty((stop - start, self.n_atoms, 3), dtype=np.float32)
        i = start
        while i < stop:
            n = self.planar_count(i, stop)
            k, a = self._locate(i)
            out[i - start:i - start + n] = self.parts[k].block(a, a + n)
            i += n
        return out

    def frame(self, i):
        k, a = self._locate(int(i))
        return self.parts[k].frame(a)

    def dimensions_of(self, i):
        k, a = self._locate(int(i))
        return self.parts[k].dimensions_of(a)


# ------------------------------------------------------------------------------------------------
# XTC (GROMACS; XDR big endian, lossy integer compression, nm)
# ------------------------------------------------------------------------------------------------
def write_xtc(path, frames, dimensions=None, precision=1000.0, dt=1.0, append=False):
    """frames: float32 [F, N, 3] in Angstrom (stored in nm at 1 / precision nm resolution, like `gmx trjconv`);
    dimensions: [a, b, c, alpha, beta, gamma] in Angstrom / degrees or None (a zero box is written)."""
    from .mdhbond import _native
    from .mdhbond.hbond import box_rows
    frames = np.asarray(frames, dtype=np.float32)
    box_nm = None if dimensions is None else (box_rows(dimensions).astype(np.float64) * 0.1).astype(np.float32)
    buf, _ = _native.xtc_encode(frames, in_scale=0.1, precision=precision, box_nm=box_nm, dt=dt)
    with open(path, "ab" if append else "wb") as f:
        f.write(buf.tobytes())


class XTCFrames:
    """Frame source over an XTC file (memory mapped): ``len``, ``frame(i)``, ``block(a, b)`` decode on the host (the
    constructor needs frame 0; everything else is a convenience), ``xtc_block(a, b)`` hands out the compressed
    records for ``hb_submit_frames_xtc`` and ``dimensions_of(i)`` the unit cell in Angstrom / degrees."""

    LENGTH_SCALE = 10.0                 # nm -> Angstrom (MDAnalysis convert_units=True)

    def __init__(self, path):
        from .mdhbond import _native
        self.path = path
        self._native = _native
        self._map = np.memmap(path, dtype=np.uint8, mode="r")
        self.offsets, self.n_atoms = _native.xtc_index(self._map)
        self.n_frames = len(self.offsets) - 1
        if self.n_frames == 0:
            raise ValueError("%s: no XTC frame records" % path)

    def __len__(self):
        return self.n_frames

    def xtc_block(self, start, stop):
        """(uint8 view of the records of frames [start, stop), int64 offsets [stop - start + 1] relative to it)."""
        a, b = int(self.offsets[start]), int(self.offsets[stop])
        return self._map[a:b], self.offsets[start:stop + 1] - a

    def xtc_count(self, start, stop):
        return stop - start

    def frame(self, i):
        return self._native.xtc_decode_host(self._map, int(self.offsets[int(i)]), self.n_atoms, self.LENGTH_SCALE)[0]

    def block(self, start, stop):
        out = np.empty((stop - start, self.n_atoms, 3), dtype=np.float32)
        for k, i in enumerate(range(start, stop)):
            out[k] = self.frame(i)
        return out

    def box_nm(self, i):
        o = int(self.offsets[int(i)])
        return np.asarray(self._map[o + 16:o + 52]).view(">f4").astype(np.float32).reshape(3, 3)

    def dimensions_of(self, i):
        """[a, b, c, alpha, beta, gamma] (Angstrom, degrees) of frame i, or None for a zero box."""
        h = self.box_nm(i).astype(np.float64) * self.LENGTH_SCALE
        la, lb, lc = (float(np.linalg.norm(h[k])) for k in range(3))
        if la == 0.0 or lb == 0.0 or lc == 0.0:
            return None
        ang = lambda u, v, lu, lv: float(np.rad2deg(np.arccos(np.clip(np.dot(u, v) / (lu * lv), -1.0, 1.0))))
        return [la, lb, lc, ang(h[1], h[2], lb, lc), ang(h[0], h[2], la, lc), ang(h[0], h[1], la, lb)]


class ConcatXTCFrames(ConcatDCDFrames):
    """Several XTC files of one system read as one trajectory; ``xtc_block`` stays inside one file."""

    def __init__(self, paths):
        self.parts = [XTCFrames(p) for p in paths]
        if not self.parts:
            raise ValueError("no trajectory files")
        self.n_atoms = self.parts[0].n_atoms
        for p, part in zip(paths, self.parts):
            if part.n_atoms != self.n_atoms:
                raise ValueError("%s has %d atoms, %s has %d" % (p, part.n_atoms, paths[0], self.n_atoms))
        self.offsets = np.concatenate([[0], np.cumsum([len(p) for p in self.parts])]).astype(np.int64)
        self.n_frames = int(self.offsets[-1])
        self.has_cell = True

    planar_block = None                  # (not a planar source; hbond.py looks for xtc_block first)

    def xtc_block(self, start, stop):
        k, a = self._locate(start)
        b = min(int(stop - self.offsets[k]), len(self.parts[k]))
        return self.parts[k].xtc_block(a, b)

    def xtc_count(self, start, stop):
        return self.planar_count(start, stop)


def universe_from_files(structure, trajectory=None):
    """``Universe`` from a PDB or PSF file and DCD or XTC trajectories (a PDB alone: its coordinates are the only frame)."""
    import os
    if str(structure).lower().endswith(".psf"):
        if trajectory is None:
            raise ValueError("%s: a PSF topology holds no coordinates, a trajectory is needed" % structure)
        topo, xyz, dims = read_psf(structure), None, None
    else:
        topo, xyz, dims = read_pdb(structure)
    if trajectory is None:
        return Universe(topo, xyz[None], dims)
    paths = list(trajectory) if isinstance(trajectory, (list, tuple)) else [trajectory]
    exts = {os.path.splitext(str(p))[1].lower() for p in paths}
    if exts == {".xtc"}:
        frames = ConcatXTCFrames(paths) if len(paths) > 1 else XTCFrames(paths[0])
    elif exts == {".dcd"}:
        frames = ConcatDCDFrames(paths) if len(paths) > 1 else DCDFrames(paths[0])
    else:
        raise ValueError("trajectories must be all .dcd or all .xtc files, got %s" % sorted(exts))
    if frames.n_atoms != topo.n_atoms:
        raise ValueError("%s has %d atoms, %s has %d" % (trajectory, frames.n_atoms, structure, topo.n_atoms))
    d0 = frames.dimensions_of(0) if frames.n_frames else None
    return Universe(topo, frames, d0 if d0 is not None else dims)
