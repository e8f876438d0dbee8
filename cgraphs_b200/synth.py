"""Synthetic solvated systems ("synthsys v1") for parity tests and benchmarks.

There is no network and the reference ships no structures, so every input is generated
locally, seeded and reproducible:

* topology: protein residues from a CHARMM-named template table (with or without hydrogens) placed
  as a compact globule, an optional slab of inert lipid-like filler atoms, TIP3 (``OH2 H1 H2``) or
  crystal (``HOH``/``O``) waters on a jittered grid at bulk density outside the solute;
* box: orthorhombic or triclinic; every atom keeps a vacuum skin from each cell face in every frame
  so that a minimum-image search and the reference's non-periodic search give the same edge set
  (SURVEY F1);
* trajectory: frame ``f`` = reference coordinates + a bounded, stationary displacement that is a
  closed-form function of ``(seed, atom, f)`` (a sum of cosines with random phase and frequency), so
  any frame can be produced independently - needed for frame sharding and > RAM trajectories.
  Hydrogens ride on their heavy atom plus a small wobble.

The five BASELINE.json configurations are exposed through :func:`config`.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

from .universe import Topology, Universe, FunctionFrames

# (name, parent index within residue, bond length) - parent -1 = residue origin.
# Hydrogens are the names starting with 'H'.
_BB = [("N", -1, 0.0), ("HN", 0, 1.0), ("CA", 0, 1.46), ("HA", 2, 1.09), ("C", 2, 1.52), ("O", 4, 1.23)]


def _t(side):
    """Template = backbone + side chain; side-chain parent indices are given relative to the full list."""
    return _BB + side


TEMPLATES = {
    "SER": _t([("CB", 2, 1.53), ("HB1", 6, 1.09), ("HB2", 6, 1.09), ("OG", 6, 1.42), ("HG1", 9, 0.97)]),
    "THR": _t([("CB", 2, 1.53), ("HB", 6, 1.09), ("OG1", 6, 1.42), ("HG1", 8, 0.97), ("CG2", 6, 1.52),
               ("HG21", 10, 1.09), ("HG22", 10, 1.09), ("HG23", 10, 1.09)]),
    "TYR": _t([("CB", 2, 1.53), ("CG", 6, 1.51), ("CD1", 7, 1.39), ("HD1", 8, 1.08), ("CE1", 8, 1.39),
               ("HE1", 10, 1.08), ("CZ", 10, 1.39), ("OH", 12, 1.37), ("HH", 13, 0.97), ("CD2", 7, 1.39),
               ("HD2", 15, 1.08), ("CE2", 15, 1.39), ("HE2", 17, 1.08)]),
    "ASP": _t([("CB", 2, 1.53), ("CG", 6, 1.52), ("OD1", 7, 1.25), ("OD2", 7, 1.25)]),
    "GLU": _t([("CB", 2, 1.53), ("CG", 6, 1.52), ("CD", 7, 1.52), ("OE1", 8, 1.25), ("OE2", 8, 1.25)]),
    "ASN": _t([("CB", 2, 1.53), ("CG", 6, 1.52), ("OD1", 7, 1.23), ("ND2", 7, 1.33), ("HD21", 9, 1.0),
               ("HD22", 9, 1.0)]),
    "GLN": _t([("CB", 2, 1.53), ("CG", 6, 1.52), ("CD", 7, 1.52), ("OE1", 8, 1.23), ("NE2", 8, 1.33),
               ("HE21", 10, 1.0), ("HE22", 10, 1.0)]),
    "LYS": _t([("CB", 2, 1.53), ("CG", 6, 1.52), ("CD", 7, 1.52), ("CE", 8, 1.52), ("NZ", 9, 1.49),
               ("HZ1", 10, 1.01), ("HZ2", 10, 1.01), ("HZ3", 10, 1.01)]),
    "ARG": _t([("CB", 2, 1.53), ("CG", 6, 1.52), ("CD", 7, 1.52), ("NE", 8, 1.46), ("HE", 9, 1.0),
               ("CZ", 9, 1.33), ("NH1", 11, 1.33), ("HH11", 12, 1.0), ("HH12", 12, 1.0), ("NH2", 11, 1.33),
               ("HH21", 15, 1.0), ("HH22", 15, 1.0)]),
    "HSD": _t([("CB", 2, 1.53), ("CG", 6, 1.50), ("ND1", 7, 1.38), ("HD1", 8, 1.0), ("CE1", 8, 1.32),
               ("HE1", 10, 1.08), ("NE2", 10, 1.32), ("CD2", 7, 1.36), ("HD2", 13, 1.08)]),
    "ALA": _t([("CB", 2, 1.53), ("HB1", 6, 1.09), ("HB2", 6, 1.09), ("HB3", 6, 1.09)]),
    "LEU": _t([("CB", 2, 1.53), ("CG", 6, 1.53), ("HG", 7, 1.09), ("CD1", 7, 1.52), ("CD2", 7, 1.52)]),
    "VAL": _t([("CB", 2, 1.53), ("HB", 6, 1.09), ("CG1", 6, 1.52), ("CG2", 6, 1.52)]),
}
POLAR = ["SER", "THR", "TYR", "ASP", "GLU", "ASN", "GLN", "LYS", "ARG", "HSD"]
APOLAR = ["ALA", "LEU", "VAL"]
RES_SPACING = 5.8      # A between residue origins in the globule
WATER_SPACING = 3.104  # A, jittered cubic grid -> 0.0334 waters / A^3
FILLER_DENSITY = 0.1   # inert atoms / A^3
SKIN = 3.0             # vacuum skin every atom keeps from each cell face
MAX_DISP = 1.2         # upper bound of |displacement| per axis (heavy 0.99 + hydrogen wobble 0.1)
N_CONF = 48            # conformers per residue type
_K_HEAVY = 4
_K_H = 2
SIGMA_HEAVY = 0.35
SIGMA_H = 0.05
FILLER_AMPLITUDE = 0.3  # fraction of the heavy-atom displacement the inert filler atoms get
SIGMA_INTRA = 0.03     # internal jitter of the heavy atoms of a (rigidly moving) protein residue


def _random_unit(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def _random_rotations(rng, n):
    """n uniformly random rotation matrices (QR of Gaussian matrices, sign-fixed)."""
    q, r = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    d = np.sign(np.diagonal(r, axis1=1, axis2=2))
    q = q * d[:, None, :]
    det = np.linalg.det(q)
    q[:, :, 0] *= det[:, None]
    return q


def _grow_conformer(rng, template):
    """Place template atoms one at a time at their bond length from the parent, rejecting
    directions that come too close to atoms already placed (hydrogens stay >= 1.65 A from every
    heavy atom but their parent, so the reference's bonded-hydrogen test `helpfunctions.py:65`
    finds exactly the template's hydrogens)."""
    n = len(template)
    xyz = np.zeros((n, 3))
    is_h = np.array([t[0][0] == "H" for t in template])
    for i, (name, parent, length) in enumerate(template):
        if parent < 0:
            continue
        others = [j for j in range(i) if j != parent]
        best, best_score = None, -np.inf
        for _ in range(200):
            p = xyz[parent] + length * _random_unit(rng, 1)[0]
            if not others:
                best = p
                break
            dist = np.linalg.norm(xyz[others] - p, axis=1)
            lim = np.where(is_h[others], 1.5, 1.65 if is_h[i] else 2.15)
            score = np.min(dist - lim)
            if score > best_score:
                best, best_score = p, score
            if score >= 0:
                break
        xyz[i] = best
    return xyz - xyz.mean(axis=0)


_CONF_CACHE = {}


def _conformers(resname, hydrogens):
    key = (resname, hydrogens)
    if key not in _CONF_CACHE:
        rng = np.random.default_rng([ord(c) for c in resname] + [7])
        tmpl = TEMPLATES[resname]
        confs = np.stack([_grow_conformer(rng, tmpl) for _ in range(N_CONF)])
        keep = [i for i, t in enumerate(tmpl) if hydrogens or t[0][0] != "H"]
        names = [tmpl[i][0] for i in keep]
        # heavy parent of each kept atom (index into kept list); heavy atoms are their own parent
        remap = {old: new for new, old in enumerate(keep)}
        parent = [remap[tmpl[i][1]] if tmpl[i][0][0] == "H" else remap[i] for i in keep]
        _CONF_CACHE[key] = (names, np.asarray(parent), confs[:, keep, :].astype(np.float64))
    return _CONF_CACHE[key]


def box_matrix(dimensions):
    """Rows = lattice vectors a, b, c from [a, b, c, alpha, beta, gamma] (degrees)."""
    a, b, c, al, be, ga = [float(x) for x in dimensions]
    al, be, ga = np.deg2rad([al, be, ga])
    ax = np.array([a, 0.0, 0.0])
    bx = np.array([b * np.cos(ga), b * np.sin(ga), 0.0])
    cx = c * np.cos(be)
    cy = c * (np.cos(al) - np.cos(be) * np.cos(ga)) / np.sin(ga)
    cz = np.sqrt(max(c * c - cx * cx - cy * cy, 0.0))
    m = np.array([ax, bx, [cx, cy, cz]])
    m[np.abs(m) < 1e-12] = 0.0
    return m


def perpendicular_widths(h):
    """Spacing between opposite faces of the cell for each lattice direction."""
    vol = abs(np.linalg.det(h))
    a, b, c = h
    return np.array([vol / np.linalg.norm(np.cross(b, c)), vol / np.linalg.norm(np.cross(c, a)),
                     vol / np.linalg.norm(np.cross(a, b))])


class SynthSystem:
    """Topology + reference coordinates + closed-form trajectory."""

    def __init__(self, topology, ref, parent, dimensions, seed, meta):
        self.topology = topology
        self.ref = np.ascontiguousarray(ref, dtype=np.float32)
        self.parent = np.asarray(parent, dtype=np.int64)
        self.dimensions = None if dimensions is None else np.asarray(dimensions, dtype=np.float32)
        self.seed = int(seed)
        self.meta = meta
        n = len(self.ref)
        rng = np.random.default_rng([self.seed, 991])
        # heavy-atom motion parameters (looked up through `parent`), hydrogen wobble parameters
        self._nu = rng.uniform(0.002, 0.05, size=(n, 3, _K_HEAVY)).astype(np.float32)
        self._phi = rng.uniform(0.0, 1.0, size=(n, 3, _K_HEAVY)).astype(np.float32)
        self._nu_h = rng.uniform(0.01, 0.2, size=(n, 3, _K_H)).astype(np.float32)
        self._phi_h = rng.uniform(0.0, 1.0, size=(n, 3, _K_H)).astype(np.float32)
        self._is_h = self.parent != np.arange(n)
        # protein residues move as rigid bodies (all heavy atoms follow the residue's first atom) plus a small internal
        # jitter, like bonded atoms in a real trajectory: independent 0.35 A displacements of bonded neighbours put
        # atoms 0.1 A apart now and then, which no force field allows (and which the XTC compression of real
        # trajectories, coded relative to the previous atom, never sees). Waters and filler atoms move as before.
        prot = np.isin(np.asarray(topology.resnames, dtype=object), list(TEMPLATES))
        first_of_res = np.zeros(n, dtype=np.int64)
        ri = np.asarray(topology.resindices)
        starts = np.nonzero(np.concatenate([[True], ri[1:] != ri[:-1]]))[0] if n else np.zeros(0, np.int64)
        first_of_res = starts[np.searchsorted(starts, np.arange(n), side="right") - 1] if n else first_of_res
        self.anchor = np.where(prot, first_of_res, np.arange(n)).astype(np.int64)
        self._w = np.where(self._is_h, 1.0, np.where(prot, SIGMA_INTRA / SIGMA_H, 0.0)).astype(np.float32)
        # inert filler atoms (a 2.15 A lattice standing in for a membrane) keep their distance: smaller amplitude
        filler = np.asarray(topology.resnames, dtype=object) == "POPC"
        self._amp = np.where(filler, np.float32(FILLER_AMPLITUDE), np.float32(1.0)).astype(np.float32)
        self._torch = {}

    @property
    def n_atoms(self):
        return len(self.ref)

    # -- numpy frames ----------------------------------------------------------------------------
    def frame(self, f):
        return self.frames(f, f + 1)[0]

    def frames(self, f0, f1, static=False):
        """float32 [f1-f0, N, 3]."""
        nf = f1 - f0
        if static:
            return np.broadcast_to(self.ref, (nf,) + self.ref.shape).copy()
        out = np.empty((nf,) + self.ref.shape, dtype=np.float32)
        two_pi = np.float32(2.0 * np.pi)
        a_heavy = np.float32(SIGMA_HEAVY * np.sqrt(2.0 / _K_HEAVY))
        a_h = np.float32(SIGMA_H * np.sqrt(2.0 / _K_H))
        for i, f in enumerate(range(f0, f1)):
            ff = np.float32(f)
            dh = (np.cos(two_pi * (self._nu * ff + self._phi))).sum(axis=2, dtype=np.float32) * a_heavy
            d = (dh * self._amp[:, None])[self.anchor[self.parent]]
            dw = (np.cos(two_pi * (self._nu_h * ff + self._phi_h))).sum(axis=2, dtype=np.float32) * a_h
            d = d + dw * self._w[:, None]
            out[i] = self.ref + d
        return out

    # -- torch frames (device-side generation for the benchmark; same formula, not bit-identical
    #    to the numpy path, so a run never mixes the two) -------------------------------------------
    def frames_torch(self, f0, f1, device, out=None):
        import torch
        key = str(device)
        if key not in self._torch:
            t = lambda a: torch.as_tensor(a, device=device)
            self._torch[key] = dict(ref=t(self.ref), parent=t(self.anchor[self.parent]), nu=t(self._nu), phi=t(self._phi),
                                    nu_h=t(self._nu_h), phi_h=t(self._phi_h), is_h=t(self._w), amp=t(self._amp))
        p = self._torch[key]
        nf = f1 - f0
        if out is None:
            out = torch.empty((nf,) + tuple(self.ref.shape), dtype=torch.float32, device=device)
        a_heavy = float(SIGMA_HEAVY * np.sqrt(2.0 / _K_HEAVY))
        a_h = float(SIGMA_H * np.sqrt(2.0 / _K_H))
        chunk = max(1, int(2e8 // (self.n_atoms * 3 * _K_HEAVY)))
        for s in range(0, nf, chunk):
            e = min(nf, s + chunk)
            ff = torch.arange(f0 + s, f0 + e, device=device, dtype=torch.float32).view(-1, 1, 1, 1)
            dh = torch.cos(2.0 * np.pi * (p["nu"][None] * ff + p["phi"][None])).sum(dim=3) * a_heavy
            d = (dh * p["amp"][None, :, None])[:, p["parent"], :]
            dw = torch.cos(2.0 * np.pi * (p["nu_h"][None] * ff + p["phi_h"][None])).sum(dim=3) * a_h
            out[s:e] = p["ref"][None] + d + dw * p["is_h"][None, :, None]
        return out

    def universe(self, n_frames=1, static=False):
        if static:
            src = FunctionFrames(n_frames, lambda i: self.ref, lambda a, b: self.frames(a, b, static=True))
        else:
            src = FunctionFrames(n_frames, self.frame, self.frames)
        return Universe(self.topology, src, self.dimensions)


def make_system(n_atoms, n_residues, seed, *, n_filler=0, hydrogens=True, water="TIP3",
                dimensions_kind="ortho", polar_fraction=0.6, n_chains=1):
    """Build a system with exactly ``n_atoms`` atoms (protein + filler + waters)."""
    rng = np.random.default_rng([seed, 17])
    # ---- protein ---------------------------------------------------------------------------------
    n_polar = int(round(n_residues * polar_fraction))
    resn = list(rng.choice(POLAR, size=n_polar)) + list(rng.choice(APOLAR, size=n_residues - n_polar))
    rng.shuffle(resn)
    rad = int(np.ceil((3.0 * n_residues / (4.0 * np.pi)) ** (1.0 / 3.0))) + 2
    g = np.arange(-rad, rad + 1)
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float64)
    lat = lat[np.argsort((lat ** 2).sum(axis=1), kind="stable")][:n_residues] * RES_SPACING
    # snake order so consecutive resids are roughly adjacent
    order = np.lexsort((lat[:, 0], lat[:, 1], lat[:, 2]))
    lat = lat[order]
    lat += rng.uniform(-0.4, 0.4, size=lat.shape)
    rots = _random_rotations(rng, n_residues)
    names, resnames, resids, segids, parents = [], [], [], [], []
    coords = []
    base = 0
    per_chain = int(np.ceil(n_residues / max(1, n_chains)))
    for r in range(n_residues):
        nm, par, confs = _conformers(resn[r], hydrogens)
        c = confs[rng.integers(N_CONF)] @ rots[r].T + lat[r]
        coords.append(c)
        names += nm
        resnames += [resn[r]] * len(nm)
        chain = r // per_chain
        resids += [(r % per_chain) + 1] * len(nm)
        segids += ["P" + chr(ord("A") + chain % 26)] * len(nm)
        parents += list(par + base)
        base += len(nm)
    prot = np.concatenate(coords) if coords else np.zeros((0, 3))
    heavy_mask = np.array([nm[0] != "H" for nm in names], dtype=bool)
    n_prot = len(prot)
    prot_radius = float(np.sqrt((prot ** 2).sum(axis=1)).max()) if n_prot else 0.0

    atoms_per_water = 3 if water == "TIP3" else 1
    n_rest = n_atoms - n_prot - n_filler
    assert n_rest >= 0, "n_atoms too small for the requested protein and filler"
    n_wat = n_rest // atoms_per_water
    n_filler += n_rest - n_wat * atoms_per_water

    # ---- box: grow the cell until filler + waters fit at their target densities -----------------------
    margin = SKIN + MAX_DISP + 1.0
    tric = dimensions_kind != "ortho"
    shape = np.sin(np.deg2rad(60.0)) if tric else 1.0
    vol = n_prot * 12.0 + n_filler / FILLER_DENSITY + n_wat * WATER_SPACING ** 3
    inner = (vol * 1.03 / shape) ** (1.0 / 3.0)
    fspacing = FILLER_DENSITY ** (-1.0 / 3.0)

    def lattice(h, hinv, fmin, spacing):
        hi = h.clip(min=0).sum(axis=0)
        lo = h.clip(max=0).sum(axis=0)
        axes = [np.arange(lo[k] + 0.5 * spacing, hi[k], spacing) for k in range(3)]
        pts = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
        fr = pts @ hinv
        return pts[np.all((fr >= fmin) & (fr <= 1.0 - fmin), axis=1)]

    for _ in range(80):
        edge = inner + 2.0 * margin / shape
        dims = [edge, edge, edge, 90.0, 90.0, 60.0 if tric else 90.0]
        h = box_matrix(dims)
        hinv = np.linalg.inv(h)
        fmin = margin / perpendicular_widths(h)
        centre = 0.5 * h.sum(axis=0)
        ptree = cKDTree(prot[heavy_mask] + centre) if n_prot else None
        filler_xyz = np.zeros((0, 3))
        fits = True
        if n_filler > 0:
            fg = lattice(h, hinv, fmin, fspacing)
            if ptree is not None:
                fg = fg[ptree.query(fg)[0] > 2.6]
            fg = fg[np.argsort(np.abs(fg[:, 2] - centre[2]), kind="stable")[:n_filler]]
            fits = len(fg) == n_filler
            filler_xyz = fg
        cand = lattice(h, hinv, fmin, WATER_SPACING)
        if ptree is not None and len(cand):
            cand = cand[ptree.query(cand)[0] > 2.4]
        if n_filler > 0 and len(cand) and len(filler_xyz):
            cand = cand[cKDTree(filler_xyz).query(cand)[0] > 2.6]
        prot_ok = True
        if n_prot:      # the protein itself must keep the skin too (small cells around large proteins)
            pf = (prot + centre) @ hinv * perpendicular_widths(h)
            pw = (1.0 - (prot + centre) @ hinv) * perpendicular_widths(h)
            prot_ok = bool(pf.min() > SKIN + MAX_DISP and pw.min() > SKIN + MAX_DISP)
        if fits and prot_ok and len(cand) >= n_wat:
            break
        inner *= 1.015
    else:
        raise RuntimeError("could not size the box")
    # keep the n_wat candidates closest to the cell centre (in the max-norm of fractional coordinates)
    fr = np.abs(cand @ hinv - 0.5).max(axis=1)
    cand = cand[np.argsort(fr, kind="stable")[:n_wat]]
    cand = cand + rng.uniform(-0.3, 0.3, size=cand.shape)
    if n_filler > 0:
        filler_xyz = filler_xyz + rng.uniform(-0.2, 0.2, size=filler_xyz.shape)

    # ---- assemble --------------------------------------------------------------------------------
    xyz = [prot + centre]
    if n_filler > 0:
        xyz.append(filler_xyz)
        per = 50
        for i in range(n_filler):
            names.append("C%d" % (i % per + 1))
            resnames.append("POPC")
            resids.append(i // per + 1)
            segids.append("MEMB")
        parents += list(range(base, base + n_filler))
        base += n_filler
    if n_wat > 0:
        if water == "TIP3":
            roh, ang = 0.9572, np.deg2rad(104.52)
            local = np.array([[0.0, 0.0, 0.0], [roh, 0.0, 0.0], [roh * np.cos(ang), roh * np.sin(ang), 0.0]])
            wr = _random_rotations(rng, n_wat)
            w = np.einsum("aj,wkj->wak", local, wr) + cand[:, None, :]
            xyz.append(w.reshape(-1, 3))
            for i in range(n_wat):
                names += ["OH2", "H1", "H2"]
                resnames += ["TIP3"] * 3
                resids += [i + 1] * 3
                segids += ["W"] * 3
                parents += [base, base, base]
                base += 3
        else:
            xyz.append(cand)
            for i in range(n_wat):
                names.append("O")
                resnames.append("HOH")
                resids.append(i + 1)
                segids.append("W")
                parents.append(base)
                base += 1
    ref = np.concatenate(xyz).astype(np.float32)
    assert len(ref) == n_atoms == len(names), (len(ref), n_atoms, len(names))
    frac = ref.astype(np.float64) @ hinv
    assert np.all(frac * perpendicular_widths(h) > SKIN + MAX_DISP - 0.3), "skin violated by the generator"
    assert np.all((1 - frac) * perpendicular_widths(h) > SKIN + MAX_DISP - 0.3), "skin violated by the generator"
    topo = Topology(names, resnames, resids, segids)
    meta = dict(n_protein_atoms=n_prot, n_filler=n_filler, n_waters=n_wat, n_residues=n_residues,
                dimensions_kind=dimensions_kind, hydrogens=hydrogens, water=water)
    return SynthSystem(topo, ref, parents, dims, seed, meta)


# ------------------------------------------------------------------------------------------------
# BASELINE.json configurations. `scale` < 1 shrinks atom and frame counts for CPU tests.
# ------------------------------------------------------------------------------------------------
CONFIGS = {
    "C1": dict(n_atoms=3000, n_residues=150, n_filler=0, hydrogens=False, water="HOH", kind="ortho",
               n_frames=1, static=True, selection="protein", check_angle=False,
               add_donors_without_hydrogen=True, method="water_around", around=3.5),
    "C2": dict(n_atoms=25000, n_residues=30, n_filler=0, hydrogens=True, water="TIP3", kind="ortho",
               n_frames=1000, static=False, selection="protein", check_angle=True,
               add_donors_without_hydrogen=False, method="water_around", around=3.5),
    "C3": dict(n_atoms=100000, n_residues=500, n_filler=20000, hydrogens=True, water="TIP3", kind="tric",
               n_frames=10000, static=False, selection="protein", check_angle=True,
               add_donors_without_hydrogen=False, method="water_around", around=3.5),
    "C4": dict(n_atoms=1000000, n_residues=5000, n_filler=200000, hydrogens=True, water="TIP3", kind="tric",
               n_frames=10000, static=False, selection="protein", check_angle=True,
               add_donors_without_hydrogen=False, method="water_around", around=3.5),
    "C5": dict(n_atoms=3000, n_residues=150, n_filler=0, hydrogens=False, water="HOH", kind="ortho",
               n_frames=1, static=True, selection="protein", check_angle=False,
               add_donors_without_hydrogen=True, method="water_around", around=3.5, n_structures=500),
}


def config(name, scale=1.0, seed=None):
    """Return ``(SynthSystem, cfg dict)`` for a BASELINE.json configuration."""
    cfg = dict(CONFIGS[name])
    if seed is None:
        seed = int(name[1:])
    n_atoms = max(300, int(round(cfg["n_atoms"] * scale)))
    n_res = max(8, int(round(cfg["n_residues"] * scale))) if name != "C2" else cfg["n_residues"]
    n_fill = int(round(cfg["n_filler"] * scale))
    cfg["n_frames"] = max(1, int(round(cfg["n_frames"] * scale))) if not cfg["static"] else 1
    sys_ = make_system(n_atoms, n_res, seed, n_filler=n_fill, hydrogens=cfg["hydrogens"], water=cfg["water"],
                       dimensions_kind=cfg["kind"])
    cfg["seed"] = seed
    return sys_, cfg


def c5_structures(n_structures=500, seed0=5000):
    """The conserved-network batch: independent static structures of 2.5k-3.5k atoms."""
    out = []
    for i in range(n_structures):
        rng = np.random.default_rng([seed0 + i, 3])
        n_atoms = int(rng.integers(2500, 3501))
        n_res = int(rng.integers(120, 181))
        out.append(make_system(n_atoms, n_res, seed0 + i, hydrogens=False, water="HOH"))
    return out


# ------------------------------------------------------------------------------------------------
# Known-answer / adversarial micro-systems (SURVEY 4, 8d "adversarial add-ons")
# ------------------------------------------------------------------------------------------------
def _step_f32(x, k):
    """x moved by k float32 ulps."""
    x = np.float32(x)
    for _ in range(abs(int(k))):
        x = np.nextafter(x, np.float32(np.inf if k > 0 else -np.inf))
    return x


def _dist2_f64(p, q):
    d = p.astype(np.float64) - q.astype(np.float64)
    return (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]


def kat_distance_system(n_pairs=1000, seed=0, r=3.5):
    """Pairs of bare water oxygens (TIP3/OH2, no hydrogens) whose separation sits within a few
    float32 ulps of the cut-off r, on both sides, plus axis-aligned pairs at exactly r.
    Use with check_angle=False, add_donors_without_hydrogen=True, selection 'resname TIP3'."""
    rng = np.random.default_rng([seed, 41])
    side = int(np.ceil(n_pairs ** (1.0 / 3.0)))
    xyz = np.zeros((2 * n_pairs, 3), dtype=np.float32)
    r2 = np.float64(r) * np.float64(r)
    for i in range(n_pairs):
        cell = np.array([i % side, (i // side) % side, i // (side * side)], dtype=np.float64)
        p1 = (cell * 12.0 + 6.0 + rng.uniform(-1, 1, 3)).astype(np.float32)
        if i % 8 == 0:   # exactly representable separation along one axis
            p1 = np.round(p1 * 2) / np.float32(2)
            p2 = p1.copy()
            ax = (i // 8) % 3
            p2[ax] = p1[ax] + np.float32(r)
            p2[ax] = _step_f32(p2[ax], (i // 24) % 3 - 1)
        else:
            u = _random_unit(rng, 1)[0]
            p2 = (p1.astype(np.float64) + r * u).astype(np.float32)
            ax = int(np.argmax(np.abs(u)))
            sgn = 1 if p2[ax] > p1[ax] else -1
            # walk outwards until just outside, then inwards to the last inside position
            for _ in range(64):
                if _dist2_f64(p1, p2) > r2:
                    break
                p2[ax] = _step_f32(p2[ax], sgn)
            for _ in range(64):
                if _dist2_f64(p1, p2) <= r2:
                    break
                p2[ax] = _step_f32(p2[ax], -sgn)
            p2[ax] = _step_f32(p2[ax], sgn * ((i % 4) - 1))     # -1, 0 (last inside), +1 (first outside), +2
        xyz[2 * i], xyz[2 * i + 1] = p1, p2
    n = 2 * n_pairs
    topo = Topology(["OH2"] * n, ["TIP3"] * n, np.arange(1, n + 1), ["W"] * n)
    return SynthSystem(topo, xyz, np.arange(n), None, seed, dict(kind="kat_distance", n_pairs=n_pairs))


def _cos_arg32(b, h, a):
    """float32 cos_arg of ANG(b, h, a) in the reference's operation order (helpfunctions.py:96-102)."""
    b, h, a = (np.asarray(t, dtype=np.float32) for t in (b, h, a))
    v1 = h - b
    v2 = a - h
    d12 = (v1[0] * v2[0] + v1[1] * v2[1]) + v1[2] * v2[2]
    d11 = (v1[0] * v1[0] + v1[1] * v1[1]) + v1[2] * v1[2]
    d22 = (v2[0] * v2[0] + v2[1] * v2[1]) + v2[2] * v2[2]
    return d12 / (np.sqrt(d11) * np.sqrt(d22))


def kat_angle_system(n_pairs=1000, seed=0, cos_threshold=0.49999994, r_oo=2.9):
    """TIP3 water pairs whose donor hydrogen makes cos_arg land within a few float32 ulps of the
    angle threshold, on both sides.  Use with check_angle=True, selection 'resname TIP3'."""
    rng = np.random.default_rng([seed, 43])
    side = int(np.ceil(n_pairs ** (1.0 / 3.0)))
    cthr = np.float32(cos_threshold)
    xyz = np.zeros((n_pairs, 2, 3, 3), dtype=np.float32)
    for i in range(n_pairs):
        cell = np.array([i % side, (i // side) % side, i // (side * side)], dtype=np.float64)
        o1 = cell * 12.0 + 6.0 + rng.uniform(-1, 1, 3)
        u = _random_unit(rng, 1)[0]
        v = np.cross(u, _random_unit(rng, 1)[0])
        v /= np.linalg.norm(v)
        o2 = o1 + r_oo * u
        o1f, o2f = o1.astype(np.float32), o2.astype(np.float32)
        # bisect the off-axis angle of the donor hydrogen so that cos_arg crosses the threshold
        lo, hi = 0.0, np.pi / 2
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            h = (o1 + 0.96 * (np.cos(mid) * u + np.sin(mid) * v)).astype(np.float32)
            if _cos_arg32(o2f, h, o1f) >= cthr:
                lo = mid
            else:
                hi = mid
        h = (o1 + 0.96 * (np.cos(lo) * u + np.sin(lo) * v)).astype(np.float32)
        ax = int(np.argmax(np.abs(v)))
        h[ax] = _step_f32(h[ax], (i % 7) - 3)
        h2 = (o1 - 0.96 * u).astype(np.float32)                      # points away from the partner
        w = np.cross(u, v)
        h3 = (o2 + 0.96 * (np.cos(0.9) * u + np.sin(0.9) * w)).astype(np.float32)
        h4 = (o2 + 0.96 * (np.cos(0.9) * u - np.sin(0.9) * w)).astype(np.float32)
        xyz[i, 0] = [o1f, h, h2]
        xyz[i, 1] = [o2f, h3, h4]
    n_w = 2 * n_pairs
    names = ["OH2", "H1", "H2"] * n_w
    resids = np.repeat(np.arange(1, n_w + 1), 3)
    topo = Topology(names, ["TIP3"] * (3 * n_w), resids, ["W"] * (3 * n_w))
    parent = np.repeat(np.arange(0, 3 * n_w, 3), 3)
    return SynthSystem(topo, xyz.reshape(-1, 3), parent, None, seed, dict(kind="kat_angle", n_pairs=n_pairs))
