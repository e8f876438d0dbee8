"""Array-backed stand-in for the slice of the MDAnalysis API that the H-bond path touches.

MDAnalysis is not installable in the build/GPU images, so trajectories reach the drop-in
``HbondAnalysis`` (and, in the parity harness, the reference's own ``mdhbond``) through this module.
The surface mirrors what the reference calls:

* ``Universe(structure[, trajectories])``, ``.select_atoms(str)``, ``.atoms``, ``.trajectory[slice]``
  (reference ``cgraphs/mdhbond/basic.py:46-50,62,68``; ``hbond.py:103,238``)
* ``AtomGroup``: ``len``, truthiness, iteration -> ``Atom``, ``[int|slice]``, ``+``, ``.atoms``,
  ``.positions`` (float32 copy)  (``helpfunctions.py:56-81``; ``network.py:55-85``)
* ``Atom``: ``name resname resid segid type index position residue``; ``Residue.atoms``.

A sliced trajectory iterator rewinds to frame 0 when exhausted (the behaviour of MDAnalysis'
``FrameIteratorSliced``), which the reference constructor depends on: it counts frames
(``basic.py:68``) and only then measures donor-hydrogen distances (``network.py:54``).

The same classes are what a user on a machine with real MDAnalysis would *not* need: the drop-in
accepts either a real ``MDAnalysis.Universe`` or a :class:`Universe` from here.
"""
from __future__ import annotations

import re
import numpy as np

# MDAnalysis' ``protein`` keyword is a fixed residue-name table; this is the subset relevant to
# CHARMM/AMBER/PDB naming that the synthetic systems and typical cgraphs inputs use.
PROTEIN_RESNAMES = frozenset("""
ALA ARG ASN ASP CYS GLN GLU GLY HIS HSD HSE HSP HID HIE HIP ILE LEU LYS MET PHE PRO SER THR TRP
TYR VAL ASH GLH LYN CYX CYM CYS1 CYS2 ACE NME NMA MSE LYR ALAD ASF CME CALA CARG CASN CASP CCYS
CGLN CGLU CGLY CHID CHIE CHIP CILE CLEU CLYS CMET CPHE CPRO CSER CTHR CTRP CTYR CVAL NALA NARG
NASN NASP NCYS NGLN NGLU NGLY NHID NHIE NHIP NILE NLEU NLYS NMET NPHE NPRO NSER NTHR NTRP NTYR
NVAL
""".split())


class Topology:
    """Per-atom string/int columns plus the residue partition."""

    def __init__(self, names, resnames, resids, segids, types=None, resindices=None):
        self.names = np.asarray(names, dtype=object)
        self.resnames = np.asarray(resnames, dtype=object)
        self.resids = np.asarray(resids, dtype=np.int64)
        self.segids = np.asarray(segids, dtype=object)
        n = len(self.names)
        if types is None:
            types = [str(nm)[0] if len(str(nm)) else "" for nm in self.names]
        self.types = np.asarray(types, dtype=object)
        if resindices is None:
            # a residue = maximal run of identical (segid, resid, resname)
            change = np.ones(n, dtype=bool)
            if n > 1:
                change[1:] = (
                    (self.resids[1:] != self.resids[:-1])
                    | (self.resnames[1:] != self.resnames[:-1])
                    | (self.segids[1:] != self.segids[:-1])
                )
            resindices = np.cumsum(change) - 1
        self.resindices = np.asarray(resindices, dtype=np.int64)
        self.n_atoms = n
        # residue -> atom indices (residues need not be contiguous, but normally are)
        order = np.argsort(self.resindices, kind="stable")
        sorted_ri = self.resindices[order]
        n_res = int(sorted_ri[-1]) + 1 if n else 0
        starts = np.searchsorted(sorted_ri, np.arange(n_res + 1))
        self._res_order = order
        self._res_starts = starts
        self.n_residues = n_res

    def residue_atoms(self, resindex):
        s, e = self._res_starts[resindex], self._res_starts[resindex + 1]
        return self._res_order[s:e]


class ArrayFrames:
    """Frame source over an in-memory float32 array ``[F, N, 3]``."""

    def __init__(self, xyz):
        xyz = np.asarray(xyz, dtype=np.float32)
        if xyz.ndim == 2:
            xyz = xyz[None]
        assert xyz.ndim == 3 and xyz.shape[2] == 3
        self.xyz = xyz

    def __len__(self):
        return self.xyz.shape[0]

    def frame(self, i):
        return self.xyz[i]

    def block(self, start, stop):
        return self.xyz[start:stop]


class FunctionFrames:
    """Frame source computed on demand: ``fn(i) -> float32 [N, 3]`` (for > RAM trajectories)."""

    def __init__(self, n_frames, fn, block_fn=None):
        self.n_frames = int(n_frames)
        self.fn = fn
        self.block_fn = block_fn

    def __len__(self):
        return self.n_frames

    def frame(self, i):
        return np.asarray(self.fn(int(i)), dtype=np.float32)

    def block(self, start, stop):
        if self.block_fn is not None:
            return np.asarray(self.block_fn(int(start), int(stop)), dtype=np.float32)
        return np.stack([self.frame(i) for i in range(start, stop)])


class Timestep:
    def __init__(self, frame, positions, dimensions):
        self.frame = frame
        self.positions = positions
        self.dimensions = dimensions


class Trajectory:
    def __init__(self, universe, frames, dimensions=None):
        self._u = universe
        self.frames = frames
        self.dimensions = dimensions  # [a, b, c, alpha, beta, gamma] or None
        self._current = 0
        self._pos = None
        self._goto(0)

    @property
    def n_frames(self):
        return len(self.frames)

    def __len__(self):
        return len(self.frames)

    def _goto(self, i):
        self._current = int(i)
        self._pos = self.frames.frame(self._current)
        self.ts = Timestep(self._current, self._pos, self.dimensions)
        return self.ts

    def rewind(self):
        return self._goto(0)

    @property
    def positions(self):
        return self._pos

    def frame_indices(self, sl):
        return range(*sl.indices(len(self.frames)))

    def __getitem__(self, item):
        if isinstance(item, slice):
            return _SlicedIterator(self, self.frame_indices(item))
        if isinstance(item, (int, np.integer)):
            n = len(self.frames)
            i = int(item)
            if i < 0:
                i += n
            if not 0 <= i < n:
                raise IndexError(item)
            return self._goto(i)
        return _SlicedIterator(self, [int(i) for i in item])

    def __iter__(self):
        return iter(_SlicedIterator(self, range(len(self.frames))))


class _SlicedIterator:
    def __init__(self, traj, indices):
        self._traj = traj
        self._indices = indices

    def __len__(self):
        return len(self._indices)

    def __iter__(self):
        for i in self._indices:
            yield self._traj._goto(i)
        self._traj.rewind()


class Residue:
    def __init__(self, universe, resindex):
        self._u = universe
        self.resindex = int(resindex)

    @property
    def atoms(self):
        return AtomGroup(self._u.topology.residue_atoms(self.resindex), self._u)


class Atom:
    __slots__ = ("_u", "index")

    def __init__(self, universe, index):
        self._u = universe
        self.index = int(index)

    @property
    def ix(self):
        return self.index

    @property
    def name(self):
        return self._u.topology.names[self.index]

    @property
    def resname(self):
        return self._u.topology.resnames[self.index]

    @property
    def resid(self):
        return int(self._u.topology.resids[self.index])

    @property
    def segid(self):
        return self._u.topology.segids[self.index]

    @property
    def type(self):
        return self._u.topology.types[self.index]

    @property
    def position(self):
        return self._u.trajectory.positions[self.index].copy()

    @property
    def residue(self):
        return Residue(self._u, self._u.topology.resindices[self.index])

    def __eq__(self, other):
        return isinstance(other, Atom) and other.index == self.index and other._u is self._u

    def __hash__(self):
        return hash(self.index)

    def __repr__(self):
        return "<Atom %d: %s of %s-%s-%d>" % (self.index, self.name, self.segid, self.resname, self.resid)


class AtomGroup:
    """Ordered list of atom indices (duplicates allowed), like ``MDAnalysis.core.groups.AtomGroup``."""

    def __init__(self, atoms=(), universe=None):
        if isinstance(atoms, AtomGroup):
            universe = atoms._u
            ix = atoms.indices
        elif isinstance(atoms, np.ndarray):
            ix = atoms.astype(np.int64, copy=False)
        else:
            atoms = list(atoms)
            if atoms and isinstance(atoms[0], Atom):
                universe = atoms[0]._u
                ix = np.fromiter((a.index for a in atoms), dtype=np.int64, count=len(atoms))
            else:
                ix = np.asarray(atoms, dtype=np.int64)
        self._u = universe
        self.indices = np.ascontiguousarray(ix, dtype=np.int64).reshape(-1)

    @property
    def ix(self):
        return self.indices

    @property
    def universe(self):
        return self._u

    @property
    def atoms(self):
        return self

    @property
    def n_atoms(self):
        return len(self.indices)

    def __len__(self):
        return len(self.indices)

    def __bool__(self):
        return len(self.indices) > 0

    def __iter__(self):
        u = self._u
        for i in self.indices:
            yield Atom(u, i)

    def __getitem__(self, item):
        if isinstance(item, (int, np.integer)):
            return Atom(self._u, self.indices[item])
        return AtomGroup(self.indices[item], self._u)

    def __add__(self, other):
        if isinstance(other, Atom):
            other_ix = np.array([other.index], dtype=np.int64)
        else:
            other_ix = AtomGroup(other, self._u).indices
        return AtomGroup(np.concatenate([self.indices, other_ix]), self._u)

    __radd__ = None

    @property
    def positions(self):
        return np.array(self._u.trajectory.positions[self.indices], dtype=np.float32)

    @property
    def names(self):
        return self._u.topology.names[self.indices]

    @property
    def resnames(self):
        return self._u.topology.resnames[self.indices]

    @property
    def resids(self):
        return self._u.topology.resids[self.indices]

    @property
    def segids(self):
        return self._u.topology.segids[self.indices]

    @property
    def types(self):
        return self._u.topology.types[self.indices]

    @property
    def resindices(self):
        return self._u.topology.resindices[self.indices]

    def select_atoms(self, sel):
        mask = _Selector(self._u.topology, self._u).evaluate(sel)
        keep = np.unique(self.indices[mask[self.indices]])
        return AtomGroup(keep, self._u)

    def __repr__(self):
        return "<AtomGroup with %d atoms>" % len(self.indices)


class Universe:
    """``Universe(topology, frames, dimensions=None)``.

    ``frames`` is a float32 array ``[F, N, 3]`` / ``[N, 3]`` or a frame source with ``__len__``,
    ``frame(i)`` and ``block(start, stop)``.
    """

    def __init__(self, topology, frames, dimensions=None):
        self.topology = topology
        if isinstance(frames, np.ndarray) or isinstance(frames, (list, tuple)):
            frames = ArrayFrames(frames)
        self.trajectory = Trajectory(self, frames, dimensions)
        self.atoms = AtomGroup(np.arange(topology.n_atoms, dtype=np.int64), self)
        self.dimensions = dimensions

    def select_atoms(self, sel):
        mask = _Selector(self.topology, self).evaluate(sel)
        return AtomGroup(np.nonzero(mask)[0], self)


# ------------------------------------------------------------------------------------------------
# selection mini-language: protein / backbone / nucleic / water / all / resname / name / segid / type / resid (a:b, a-b,
# a to b) / index / bynum / byres / around R <sel> (first frame) / and / or / not / ()
# precedence as in MDAnalysis: not > (and = or, left to right) > byres = around; several values after a keyword are OR-ed
# ------------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"\(|\)|[^\s()]+")
_KEYWORDS = {"and", "or", "not", "protein", "all", "resname", "name", "segid", "type", "resid", "resnum", "(", ")",
             "backbone", "nucleic", "water", "index", "bynum", "byres", "around", "chainID", "chainid"}
# MDAnalysis' residue-name tables (core/selection.py): water and nucleic acids
WATER_RESNAMES = {"H2O", "HOH", "OH2", "HHO", "OHH", "TIP", "T3P", "T4P", "T5P", "SOL", "WAT", "TIP2", "TIP3", "TIP4"}
NUCLEIC_RESNAMES = {"ADE", "URA", "CYT", "GUA", "THY", "DA", "DC", "DG", "DT", "RA", "RU", "RG", "RC", "A", "T", "U", "C", "G",
                    "DA5", "DC5", "DG5", "DT5", "DA3", "DC3", "DG3", "DT3", "RA5", "RU5", "RG5", "RC5", "RA3", "RU3", "RG3", "RC3"}
BACKBONE_NAMES = {"N", "CA", "C", "O"}


class _Selector:
    def __init__(self, topology, universe=None):
        self.t = topology
        self.u = universe

    def evaluate(self, sel):
        self.tokens = _TOKEN.findall(sel)
        self.pos = 0
        if not self.tokens:
            raise ValueError("empty selection")
        mask = self._expr(0)
        if self.pos != len(self.tokens):
            raise ValueError("unexpected token %r in selection %r" % (self.tokens[self.pos], sel))
        return mask

    def _peek(self):
        return self.tokens[self.pos] if self.pos < len(self.tokens) else None

    def _next(self):
        tok = self._peek()
        self.pos += 1
        return tok

    # MDAnalysis' SelectionParser (core/selection.py) is a precedence-climbing parser: `and` and `or` share one
    # precedence (3) and associate to the left, `not` (5) binds tighter, `byres` and `around` (1) take the whole
    # expression that follows them:  A or B and C = (A or B) and C;  around 3.5 protein and name CA =
    # around 3.5 (protein and name CA);  not A and B = (not A) and B.
    _PREC = {"and": 3, "or": 3}

    def _expr(self, p=0):
        m = self._subexp()
        while self._peek() in self._PREC and self._PREC[self._peek()] >= p:
            op = self._next()
            rhs = self._expr(self._PREC[op] + 1)
            m = (m & rhs) if op == "and" else (m | rhs)
        return m

    def _subexp(self):
        if self._peek() == "not":
            self._next()
            return ~self._expr(5)
        if self._peek() == "byres":             # every atom of the residues touched by the following expression
            self._next()
            inner = self._expr(1)
            hit = np.zeros(self.t.n_residues, dtype=bool)
            hit[self.t.resindices[inner]] = True
            return hit[self.t.resindices]
        if self._peek() == "around":            # atoms within R of the following expression (first frame), not the expression itself
            self._next()
            radius = float(self._next())
            inner = self._expr(1)
            if self.u is None:
                raise ValueError("'around' needs coordinates")
            from scipy.spatial import cKDTree
            pos = np.asarray(self.u.trajectory.frames.frame(0), dtype=np.float32)
            m = np.zeros(self.t.n_atoms, dtype=bool)
            if inner.any():
                near = cKDTree(pos).query_ball_tree(cKDTree(pos[inner]), radius)
                m = np.fromiter((len(x) > 0 for x in near), dtype=bool, count=self.t.n_atoms)
            return m & ~inner
        return self._atom()

    def _values(self):
        vals = []
        while self._peek() is not None and self._peek() not in _KEYWORDS:
            vals.append(self._next())
        if not vals:
            raise ValueError("selection keyword without value")
        return vals

    @staticmethod
    def _match(column, vals):
        m = np.zeros(len(column), dtype=bool)
        for v in vals:
            if "*" in v or "?" in v:
                rx = re.compile("^" + re.escape(v).replace(r"\*", ".*").replace(r"\?", ".") + "$")
                m |= np.fromiter((rx.match(str(c)) is not None for c in column), dtype=bool, count=len(column))
            else:
                m |= column == v
        return m

    def _atom(self):
        tok = self._next()
        t = self.t
        if tok is None:
            raise ValueError("unexpected end of selection")
        if tok == "(":
            m = self._expr(0)
            if self._next() != ")":
                raise ValueError("missing ) in selection")
            return m
        if tok == "all":
            return np.ones(t.n_atoms, dtype=bool)
        if tok == "protein":
            return np.fromiter((r in PROTEIN_RESNAMES for r in t.resnames), dtype=bool, count=t.n_atoms)
        if tok == "backbone":
            prot = np.fromiter((r in PROTEIN_RESNAMES for r in t.resnames), dtype=bool, count=t.n_atoms)
            return prot & np.fromiter((n in BACKBONE_NAMES for n in t.names), dtype=bool, count=t.n_atoms)
        if tok == "nucleic":
            return np.fromiter((r in NUCLEIC_RESNAMES for r in t.resnames), dtype=bool, count=t.n_atoms)
        if tok == "water":
            return np.fromiter((r in WATER_RESNAMES for r in t.resnames), dtype=bool, count=t.n_atoms)
        if tok in ("index", "bynum"):
            m = np.zeros(t.n_atoms, dtype=bool)
            shift = 1 if tok == "bynum" else 0          # bynum is 1-based, both ranges are inclusive
            for v in self._values():
                mm = re.match(r"^(\d+)(?:[:\-](\d+))?$", v)
                if not mm:
                    raise ValueError("bad %s range %r" % (tok, v))
                lo = int(mm.group(1)) - shift
                hi = (int(mm.group(2)) if mm.group(2) is not None else int(mm.group(1))) - shift
                m[max(lo, 0):hi + 1] = True
            return m
        if tok in ("chainID", "chainid"):               # (the built-in readers keep the chain id in segids)
            return self._match(t.segids, self._values())
        if tok == "resname":
            return self._match(t.resnames, self._values())
        if tok == "name":
            return self._match(t.names, self._values())
        if tok == "segid":
            return self._match(t.segids, self._values())
        if tok == "type":
            return self._match(t.types, self._values())
        if tok in ("resid", "resnum"):
            m = np.zeros(t.n_atoms, dtype=bool)
            vals = self._values()
            k = 0
            while k < len(vals):
                v = vals[k]
                if k + 2 < len(vals) and vals[k + 1] == "to":      # 'resid 1 to 5'
                    v = v + ":" + vals[k + 2]
                    k += 2
                k += 1
                mm = re.match(r"^(-?\d+)(?:[:\-](-?\d+))?$", v)
                if not mm:
                    raise ValueError("bad resid range %r" % v)
                lo = int(mm.group(1))
                hi = int(mm.group(2)) if mm.group(2) is not None else lo
                m |= (t.resids >= lo) & (t.resids <= hi)
            return m
        raise ValueError("unknown selection token %r" % tok)
