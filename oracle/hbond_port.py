"""TEST INFRASTRUCTURE - CPU restatement ("port") of the reference H-bond hot path.

This module is the parity oracle for the CUDA path and the `cpu_baseline` arm of `bench.py`.  It is
NOT part of the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it.  The product (`cgraphs_b200.mdhbond`) never does and fails
loudly without its CUDA library.

It restates, from scratch, the algorithm of evabertalan/cgraphs `cgraphs/mdhbond`:

* construction           basic.py:34-69, network.py:40-97, helpfunctions.py:30-34,44-83,107-117
* `hbonds_in_selection`  hbond.py:96-135
* `hbonds_in_selection_and_water_around`  hbond.py:232-289
* angle criterion        helpfunctions.py:86-104,137-159
* result publication     basic.py:94-98, helpfunctions.py:120-134

The arithmetic of the reference lives in third-party code that is not under /root/reference:
scipy.spatial.cKDTree (unpinned by the reference; scipy 1.18.1 in this image) for the distance
queries and numpy (2.3.5 here) float32 ufuncs for the angle.  The reference has no tests, fixtures
or golden vectors, so parity is pinned by executing the reference itself in the build
container: `tests/test_oracle_vs_reference.py` runs this port against the live reference
(`oracle/ref_harness.py`) and `tests/golden/*.json` hold outputs of the live reference generated
by `tests/golden/make_golden.py`.

Two distance back-ends are provided and tested equal: 'kdtree' issues the same cKDTree queries
as the reference; 'brute' evaluates the predicate
    DIST(p, q, r):  fl64(fl64(dx*dx + dy*dy) + dz*dz) <= fl64(r*r)   on float32 inputs widened to float64
directly, which is what the CUDA kernels implement.

Deliberate divergence (documented in DESIGN.md): frames with no candidate pair / no local water make
the reference raise IndexError (hbond.py:110,254,261); the port, like the product, treats them as
frames without bonds unless `strict=True`.
"""
from __future__ import annotations

import itertools
import numpy as np
import networkx as nx
from scipy.spatial import cKDTree

DONOR_NAMES = {'OH2', 'OW', 'NE', 'NH1', 'NH2', 'ND2', 'SG', 'NE2', 'ND1', 'NZ', 'OG', 'OG1', 'NE1', 'OH',
               'OE1', 'OE2', 'N16', 'OD1', 'OD2'}                      # helpfunctions.py:31
ACCEPTOR_NAMES = {'OH2', 'OW', 'OD1', 'OD2', 'SG', 'OE1', 'OE2', 'ND1', 'NE2', 'SD', 'OG', 'OG1', 'OH'}  # :32
WATER_DEFINITION = '(resname TIP3 and name OH2) or (resname HOH and name O) or (resname TIP4 and name OH2)'  # :34


def _r_covalent(first_letter):                                          # helpfunctions.py:30
    return {'N': 1.31, 'O': 1.31, 'P': 1.58, 'S': 1.55}.get(first_letter, 1.5)


def _id(topo, i, detailed):                                             # helpfunctions.py:107-117
    s = str(topo.segids[i]) + '-' + str(topo.resnames[i]) + '-' + str(int(topo.resids[i]))
    return s + '-' + str(topo.names[i]) if detailed else s


class PortTopology:
    """Everything the reference constructor derives, as plain arrays/lists."""

    def __init__(self, universe, selection, *, distance=3.5, cut_angle=60., start=None, stop=None, step=1,
                 additional_donors=(), additional_acceptors=(), exclude_donors=(), exclude_acceptors=(),
                 check_angle=True, residuewise=True, add_donors_without_hydrogen=False):
        if selection is None:
            raise AssertionError('No selection string.')
        if universe is None:
            raise AssertionError('No structure file path.')
        topo = universe.topology
        self.universe = universe
        self.frame_slice = slice(start if isinstance(start, int) else None,
                                 stop if isinstance(stop, int) else None, step)
        self.frame_indices = list(universe.trajectory.frame_indices(self.frame_slice))
        self.nb_frames = len(self.frame_indices)
        sel = universe.select_atoms(selection).indices
        if len(sel) == 0:
            raise AssertionError('No atoms match the selection')
        self.water_atoms = universe.select_atoms(WATER_DEFINITION).indices
        donor_names = (DONOR_NAMES | set(additional_donors)) - set(exclude_donors)
        acceptor_names = (ACCEPTOR_NAMES | set(additional_acceptors)) - set(exclude_acceptors)
        # hydrogen detection uses the coordinates of frame 0 (the frame-count loop rewinds, A.1 step 5)
        pos0 = universe.trajectory.frames.frame(0)

        donors, acceptors, hydrogens, donor2h = [], [], [], []
        for a in sel:
            name = str(topo.names[a])
            polar = (len(name) > 1) and name[0] in ('N', 'O')
            if name in donor_names or polar:
                if add_donors_without_hydrogen:
                    donors.append(a)
                else:
                    rc = np.float32(_r_covalent(name[0]))
                    found = 0
                    for o in topo.residue_atoms(topo.resindices[a]):
                        on = str(topo.names[o])
                        if on[0] == 'H' or on[:2] in ('1H', '2H', '3H') or topo.types[o] == 'H':
                            d = np.sqrt(((pos0[a] - pos0[o]) ** 2).sum())   # float32 throughout
                            if d < rc:
                                hydrogens.append(o)
                                found += 1
                    if found:
                        donor2h.append(list(range(len(hydrogens) - found, len(hydrogens))))
                        donors.append(a)
            if name in acceptor_names or polar:
                acceptors.append(a)
        if len(donors) + len(acceptors) == 0:
            raise AssertionError('neither donors nor acceptors in the selection')
        self.nD, self.nA, self.nW = len(donors), len(acceptors), len(self.water_atoms)
        self.donors = np.asarray(donors, dtype=np.int64)
        self.acceptors = np.asarray(acceptors, dtype=np.int64)
        self.entry_atom = np.concatenate([self.donors, self.acceptors, self.water_atoms]).astype(np.int64)
        self.first_water = self.nD + self.nA
        n_sel_h = len(hydrogens)
        water_h = []
        if not add_donors_without_hydrogen:
            for w in self.water_atoms:
                water_h.extend(topo.residue_atoms(topo.resindices[w])[1:])
        if not hydrogens and not water_h and check_angle:
            raise AssertionError('There are no possible hbond donors in the selection and no water. Since '
                                 'check_angle is True, hydrogen is needed for the calculations!')
        self.h_atom = np.asarray(hydrogens + water_h, dtype=np.int64)
        self.heavy2hydrogen = (donor2h + [[] for _ in acceptors]
                               + [[n_sel_h + i, n_sel_h + i + 1] for i in range(0, len(water_h), 2)])
        nS = self.first_water
        self.water_ids = [_id(topo, i, False) for i in self.water_atoms]
        self.water_ids_atomwise = [_id(topo, i, True) for i in self.water_atoms]
        self.all_ids = [_id(topo, i, not residuewise) for i in self.entry_atom[:nS]] + self.water_ids
        self.all_ids_atomwise = [_id(topo, i, True) for i in self.entry_atom[:nS]] + self.water_ids_atomwise
        self.resids = np.array([int(s.split('-')[2]) for s in self.all_ids], dtype=np.int64)
        self.distance, self.cut_angle = distance, cut_angle
        self.check_angle, self.residuewise = check_angle, residuewise


def angle_deg(p1, p2, p3):
    """helpfunctions.py:86-104 restated with explicit float32 operation order."""
    v1 = p2 - p1
    v2 = p3 - p2
    d12 = (v1[:, 0] * v2[:, 0] + v1[:, 1] * v2[:, 1]) + v1[:, 2] * v2[:, 2]
    d11 = (v1[:, 0] * v1[:, 0] + v1[:, 1] * v1[:, 1]) + v1[:, 2] * v1[:, 2]
    d22 = (v2[:, 0] * v2[:, 0] + v2[:, 1] * v2[:, 1]) + v2[:, 2] * v2[:, 2]
    with np.errstate(all='ignore'):
        c = d12 / (np.sqrt(d11) * np.sqrt(d22))
        c[c < -1.0] = -1.0
        return np.rad2deg(np.arccos(c))


def cos_arg(p1, p2, p3):
    """The float32 `cos_arg` of helpfunctions.py:96-102 (before the clamp)."""
    v1 = p2 - p1
    v2 = p3 - p2
    d12 = (v1[:, 0] * v2[:, 0] + v1[:, 1] * v2[:, 1]) + v1[:, 2] * v2[:, 2]
    d11 = (v1[:, 0] * v1[:, 0] + v1[:, 1] * v1[:, 1]) + v1[:, 2] * v1[:, 2]
    d22 = (v2[:, 0] * v2[:, 0] + v2[:, 1] * v2[:, 1]) + v2[:, 2] * v2[:, 2]
    with np.errstate(all='ignore'):
        return d12 / (np.sqrt(d11) * np.sqrt(d22))


def _flatten(lists):
    lens = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
    i = np.repeat(np.arange(len(lists), dtype=np.int64), lens)
    j = np.fromiter(itertools.chain.from_iterable(lists), dtype=np.int64, count=int(lens.sum()))
    return i, j


def _dist2_f64(a, b):
    d = a.astype(np.float64)[:, None, :] - b.astype(np.float64)[None, :, :]
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def ball_pairs(a, b, r, backend='kdtree'):
    """All (i, j) with DIST(a_i, b_j, r)."""
    if len(a) == 0 or len(b) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    if backend == 'kdtree':
        return _flatten(cKDTree(a).query_ball_tree(cKDTree(b), r))
    r2 = np.float64(r) * np.float64(r)
    out_i, out_j = [], []
    step = max(1, int(4e6 // max(1, len(b))))
    for s in range(0, len(a), step):
        i, j = np.nonzero(_dist2_f64(a[s:s + step], b) <= r2)
        out_i.append(i + s)
        out_j.append(j)
    return np.concatenate(out_i), np.concatenate(out_j)


def self_pairs(a, r, backend='kdtree'):
    """All i < j with DIST(a_i, a_j, r)."""
    if len(a) < 2:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    if backend == 'kdtree':
        p = cKDTree(a).query_pairs(r, output_type='ndarray')
        return p[:, 0].astype(np.int64), p[:, 1].astype(np.int64)
    i, j = ball_pairs(a, a, r, 'brute')
    keep = i < j
    return i[keep], j[keep]


# ------------------------------------------------------------------------------------------------
# Minimum-image extension (NOT in the reference, which applies no periodic boundary conditions - SURVEY F1).
# Definition shared with the kernels: lattice vectors are lower-triangular rows a, b, c (float32);
# a difference vector is reduced along c, then b, then a with n = rint(d_k / L_k) and individually rounded
# multiply / subtract - in float64 on the float64 differences for DIST, in float32 for the angle vectors. When
# nothing wraps (n = 0) the arithmetic is exactly the reference's.
# ------------------------------------------------------------------------------------------------
def box_rows(dimensions):
    """[a, b, c, alpha, beta, gamma] -> float32 3x3, rows = lattice vectors (MDAnalysis triclinic_vectors layout)."""
    a, b, c, al, be, ga = [float(x) for x in dimensions]
    al, be, ga = np.deg2rad([al, be, ga])
    m = np.zeros((3, 3), dtype=np.float64)
    m[0, 0] = a
    m[1, 0] = b * np.cos(ga)
    m[1, 1] = b * np.sin(ga)
    m[2, 0] = c * np.cos(be)
    m[2, 1] = c * (np.cos(al) - np.cos(be) * np.cos(ga)) / np.sin(ga)
    m[2, 2] = np.sqrt(max(c * c - m[2, 0] ** 2 - m[2, 1] ** 2, 0.0))
    m[np.abs(m) < 1e-12] = 0.0
    return m.astype(np.float32)


def _reduce(dx, dy, dz, box, dtype):
    ax, bx, by, cx, cy, cz = [dtype(box[i, j]) for i, j in ((0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2))]
    n = np.rint(dz / cz)
    dz = dz - n * cz
    dy = dy - n * cy
    dx = dx - n * cx
    n = np.rint(dy / by)
    dy = dy - n * by
    dx = dx - n * bx
    n = np.rint(dx / ax)
    dx = dx - n * ax
    return dx, dy, dz


def dist2_pbc(pa, pb, box):
    """float64 squared minimum-image distance of float32 points pa[k], pb[k] (same operation order as dist_le_pbc)."""
    d = pa.astype(np.float64) - pb.astype(np.float64)
    dx, dy, dz = _reduce(d[:, 0].copy(), d[:, 1].copy(), d[:, 2].copy(), box, np.float64)
    return (dx * dx + dy * dy) + dz * dz


def min_image32(v, box):
    vx, vy, vz = _reduce(v[:, 0].copy(), v[:, 1].copy(), v[:, 2].copy(), box, np.float32)
    return np.stack([vx, vy, vz], axis=1)


def ball_pairs_pbc(a, b, r, box):
    """All (i, j) with minimum-image DIST(a_i, b_j, r): candidates from a KD-tree over the 27 images of b."""
    if len(a) == 0 or len(b) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    h = box.astype(np.float64)
    shifts = np.array([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)], dtype=np.float64) @ h
    # bring everything into one cell first so that +-1 images suffice for the candidate search
    def wrap(p):
        f = p.astype(np.float64) @ np.linalg.inv(h)
        return (f - np.floor(f)) @ h
    aw, bw = wrap(a), wrap(b)
    images = (bw[None, :, :] + shifts[:, None, :]).reshape(-1, 3)
    lists = cKDTree(aw).query_ball_tree(cKDTree(images), float(r) * 1.001 + 1e-3)
    i, j = _flatten(lists)
    j = j % len(b)
    if len(i):
        ij = np.unique(np.stack([i, j], axis=1), axis=0)
        i, j = ij[:, 0], ij[:, 1]
        keep = dist2_pbc(a[i], b[j], box) <= np.float64(r) * np.float64(r)
        i, j = i[keep], j[keep]
    return i, j


def self_pairs_pbc(a, r, box):
    i, j = ball_pairs_pbc(a, a, r, box)
    keep = i < j
    return i[keep], j[keep]


def _angle_filter(pt, e0, e1, coords, hcoords, cut_angle, cos_threshold=None, box=None):
    """helpfunctions.py:137-159: keep a pair if any hydrogen of either partner passes the angle test."""
    if len(e0) == 0:
        return e0, e1
    h2h = pt.heavy2hydrogen
    cnt = np.fromiter((len(h2h[a]) + len(h2h[b]) for a, b in zip(e0, e1)), dtype=np.int64, count=len(e0))
    pair_of = np.repeat(np.arange(len(e0)), cnt)
    hidx = np.fromiter(itertools.chain.from_iterable(
        itertools.chain(h2h[a], h2h[b]) for a, b in zip(e0, e1)), dtype=np.int64, count=int(cnt.sum()))
    pa = coords[e0[pair_of]]
    pb = coords[e1[pair_of]]
    ph = hcoords[hidx]
    if box is not None:       # minimum-image vectors: v1 = h - b, v2 = a - h, each reduced in float32
        v1 = min_image32(ph - pb, box)
        v2 = min_image32(pa - ph, box)
        d12 = (v1[:, 0] * v2[:, 0] + v1[:, 1] * v2[:, 1]) + v1[:, 2] * v2[:, 2]
        d11 = (v1[:, 0] * v1[:, 0] + v1[:, 1] * v1[:, 1]) + v1[:, 2] * v1[:, 2]
        d22 = (v2[:, 0] * v2[:, 0] + v2[:, 1] * v2[:, 1]) + v2[:, 2] * v2[:, 2]
        with np.errstate(all='ignore'):
            c = d12 / (np.sqrt(d11) * np.sqrt(d22))
            if cos_threshold is None:
                cc = c.copy()
                cc[cc < -1.0] = -1.0
                ok = np.rad2deg(np.arccos(cc)) <= np.float32(cut_angle)
            else:
                ok = (np.where(c < -1.0, np.float32(-1.0), c) >= np.float32(cos_threshold)) & (c <= np.float32(1.0))
    elif cos_threshold is None:
        ok = angle_deg(pb, ph, pa) <= np.float32(cut_angle)
    else:  # the threshold form the kernels use; tested equal to the line above
        c = cos_arg(pb, ph, pa)
        with np.errstate(all='ignore'):
            ok = (c >= np.float32(cos_threshold)) & (c <= np.float32(1.0))
    keep = np.zeros(len(e0), dtype=bool)
    keep[pair_of[ok]] = True
    return e0[keep], e1[keep]


def _frame_keys(pt, e0, e1, ids):
    """hbond.py:119-127 / 271-281: orientation by resid, optional same-residue skip; returns key strings."""
    x = np.minimum(e0, e1)
    y = np.maximum(e0, e1)
    check = pt.resids[x] < pt.resids[y]
    first = np.where(check, x, y)
    second = np.where(check, y, x)
    keys = set()
    for f, s in set(zip(first.tolist(), second.tolist())):
        a, b = ids[f], ids[s]
        if not pt.check_angle and a.split('-')[:3] == b.split('-')[:3]:
            continue
        keys.add(a + ':' + b)
    return keys


def _backbone_flags(pt):
    return np.array([(s.split('-')[3] in ('O', 'N')) for s in pt.all_ids_atomwise], dtype=bool)


def frame_pairs(pt, pos, method, around=None, exclude_backbone_backbone=True, backend='kdtree',
                cos_threshold=None, strict=False, box=None):
    """Entry pairs (col0, col1) that are H-bonds in one frame, before key building.
    box (float32 3x3 lattice rows) switches to the minimum-image extension."""
    nD, nA, fw = pt.nD, pt.nA, pt.first_water
    if box is not None:
        box = np.asarray(box, dtype=np.float32)
        _bp = lambda a, b, r, backend: ball_pairs_pbc(a, b, r, box)
        _sp = lambda a, r, backend: self_pairs_pbc(a, r, box)
    else:
        _bp, _sp = ball_pairs, self_pairs
    coords_sel = pos[pt.entry_atom[:fw]]
    e0 = np.zeros(0, np.int64)
    e1 = np.zeros(0, np.int64)
    if nD > 0 and nA > 0:
        ia, jd = _bp(pos[pt.acceptors], pos[pt.donors], pt.distance, backend)
        if strict and len(ia) == 0:
            raise IndexError('reference raises on frames without donor-acceptor pairs')
        e0, e1 = ia + nD, jd
        if exclude_backbone_backbone and len(e0):
            bb = _backbone_flags(pt)
            keep = ~(bb[e0] & bb[e1])
            e0, e1 = e0[keep], e1[keep]
    if method == 'in_selection':
        coords = coords_sel
    else:
        wpos = pos[pt.water_atoms]
        _, wj = _bp(coords_sel, wpos, float(around), backend)
        local = np.unique(wj)
        if strict and len(local) == 0:
            raise IndexError('reference raises on frames without local water')
        lpos = wpos[local]
        si, lj = _bp(coords_sel, lpos, pt.distance, backend)
        wi, wj2 = _sp(lpos, pt.distance, backend)
        # candidate order of the reference: DA ++ WW ++ SW (irrelevant for the result)
        e0 = np.concatenate([e0, local[wi] + fw, si])
        e1 = np.concatenate([e1, local[wj2] + fw, local[lj] + fw])
        coords = np.vstack([coords_sel, wpos])
    if pt.check_angle:
        e0, e1 = _angle_filter(pt, e0, e1, coords, pos[pt.h_atom], pt.cut_angle, cos_threshold, box)
    return e0, e1


class PortResult:
    def __init__(self, pt, results):
        self.topology = pt
        self.initial_results = results
        self.filtered_results = results
        self.nb_frames = pt.nb_frames
        self.residuewise = pt.residuewise
        self.initial_graph = dict2graph(results, pt.residuewise)
        self.filtered_graph = self.initial_graph


def dict2graph(bonds, residuewise=True):
    """helpfunctions.py:120-134."""
    g = nx.Graph()
    cut = 3 if residuewise else 4
    edges = []
    for bond in bonds:
        a, b = bond.split(':')
        a = '-'.join(a.split('-')[:cut])
        b = '-'.join(b.split('-')[:cut])
        if int(a.split('-')[2]) > int(b.split('-')[2]):
            a, b = b, a
        if a != b:
            edges.append((a, b))
    g.add_edges_from(edges)
    return g


def key_ids(pt, method):
    """hbond.py:122 (always _all_ids) / hbond.py:276-277 (_all_ids or _all_ids_atomwise)."""
    if method == 'in_selection' or pt.residuewise:
        return pt.all_ids
    return pt.all_ids_atomwise


def run_block(pt, xyz, method, around=3.5, exclude_backbone_backbone=True, backend='kdtree',
              cos_threshold=None, strict=False, results=None, col0=0, nb_frames=None, box=None):
    """Process the frames xyz[k] (float32 [K, N, 3]) into `results` columns col0 + k; returns results.
    box: None, one 3x3 lattice for all frames or an array [K, 3, 3] (minimum-image extension)."""
    ids = key_ids(pt, method)
    if results is None:
        results = {}
    if nb_frames is None:
        nb_frames = pt.nb_frames
    for k in range(len(xyz)):
        pos = np.asarray(xyz[k], dtype=np.float32)
        bk = None if box is None else (np.asarray(box)[k] if np.asarray(box).ndim == 3 else np.asarray(box))
        e0, e1 = frame_pairs(pt, pos, method, around, exclude_backbone_backbone, backend, cos_threshold, strict, bk)
        for key in _frame_keys(pt, e0, e1, ids):
            row = results.get(key)
            if row is None:
                row = results[key] = np.zeros(nb_frames, dtype=bool)
            row[col0 + k] = True
    return results


def run(universe, selection, method='in_selection', around=3.5, exclude_backbone_backbone=True,
        backend='kdtree', cos_threshold=None, strict=False, frame_range=None, box=None, **ctor):
    """Run the port; returns a PortResult whose `initial_results` is `dict[str, bool[nb_frames]]`."""
    pt = PortTopology(universe, selection, **ctor)
    if method == 'in_selection':
        if pt.nA == 0:
            raise AssertionError('No acceptors in the selection')
        if pt.nD == 0:
            raise AssertionError('No donors in the selection')
    results = {}
    frames = universe.trajectory.frames
    for col, f in enumerate(pt.frame_indices):
        if frame_range is not None and not (frame_range[0] <= col < frame_range[1]):
            continue
        run_block(pt, [frames.frame(f)], method, around, exclude_backbone_backbone, backend, cos_threshold,
                  strict, results, col, box=box)
    return PortResult(pt, results)


def run_ions(universe, selection, ions, include_water=True, box=None, **ctor):
    """HbondAnalysis.set_add_ion_interactions (hbond.py:57-94; ion group basic.py:53-60): contacts within `distance`
    of every selection entry (and every water of the water group) with the ions, no angle criterion.
    Returns the result dict the reference merges into the H-bond results."""
    pt = PortTopology(universe, selection, **ctor)
    topo = universe.topology
    ion_atoms = universe.select_atoms(' or '.join('name %s' % i for i in ions)).indices
    detailed = not pt.residuewise
    ion_ids = [_id(topo, i, detailed) for i in ion_atoms]
    water_ids = pt.water_ids if pt.residuewise else pt.water_ids_atomwise
    fw = pt.first_water
    results = {}
    frames = universe.trajectory.frames
    bp = (lambda a, b, r: ball_pairs_pbc(a, b, r, np.asarray(box, dtype=np.float32))) if box is not None else \
         (lambda a, b, r: ball_pairs(a, b, r))
    for col, f in enumerate(pt.frame_indices):
        pos = np.asarray(frames.frame(f), dtype=np.float32)
        ion_pos = pos[ion_atoms]
        keys = set()
        i, j = bp(pos[pt.entry_atom[:fw]], ion_pos, pt.distance)
        keys.update(pt.all_ids[a] + ':' + ion_ids[b] for a, b in zip(i.tolist(), j.tolist()))
        if include_water and len(pt.water_atoms):
            i, j = bp(pos[pt.water_atoms], ion_pos, pt.distance)
            keys.update(water_ids[a] + ':' + ion_ids[b] for a, b in zip(i.tolist(), j.tolist()))
        for key in keys:
            row = results.get(key)
            if row is None:
                row = results[key] = np.zeros(pt.nb_frames, dtype=bool)
            row[col] = True
    return results


# ------------------------------------------------------------------------------------------------
# WireAnalysis.set_water_wires (waterwire.py:191-317) - neighbour stage :200-248,
# per-frame breadth-first search :250-295, direct bonds :297-309, publication :312-314.
# ------------------------------------------------------------------------------------------------
def wire_node_of_entry(pt):
    """`da_trans` of waterwire.py:60-65: every selection entry is represented by the FIRST entry that carries the
    same residue-level id (MDA_info_list without detailed_info, helpfunctions.py:107-117)."""
    topo = pt.universe.topology
    first = {}
    out = np.empty(pt.first_water, dtype=np.int64)
    for e in range(pt.first_water):
        out[e] = first.setdefault(_id(topo, pt.entry_atom[e], False), e)
    return out


def hull_water_labels(pt, pos, max_water, backend='kdtree'):
    """waterwire.py:212-221 with `water_in_convex_hull` set. Returns (true, label): index arrays into the water group.
    `true[j]` is the j-th local water that lies inside the Delaunay triangulation of the selection coordinates - its
    POSITION enters the distance searches (waterwire.py:221,223,226,228). The reference does not filter
    `local_water_index` along with the coordinates, so the j-th point of those searches is NAMED
    `local_water_index[j]` = `label[j]`, the j-th local water of the unfiltered list (waterwire.py:226,228): every pair
    the searches find carries the labels, and the angle criterion is evaluated with the labelled waters' coordinates
    and hydrogens (waterwire.py:236-239). Restated as it is - bit-exact results need the same labelling."""
    from scipy.spatial import Delaunay
    fw = pt.first_water
    sel = pos[pt.entry_atom[:fw]]
    wpos = pos[pt.water_atoms]
    around = float(max_water + 1) * pt.distance / 2.
    _, wj = ball_pairs(sel, wpos, around, backend)
    local = np.unique(wj)
    if len(local) == 0:
        return local, local
    inside = Delaunay(sel).find_simplex(wpos[local]) != -1
    true = local[inside]
    return true, local[:len(true)]


def wire_frame_edges(pt, pos, max_water, backend='kdtree', cos_threshold=None, box=None, water_in_convex_hull=False):
    """Neighbour stage of one frame (waterwire.py:200-248): the three H-bond pair classes, local waters taken within
    (max_water + 1) * distance / 2 of the selection, no backbone filter. Returns entry pairs
    (direct [donor entry, acceptor entry], selection-water [entry, water entry], water-water)."""
    around = float(max_water + 1) * pt.distance / 2.
    fw = pt.first_water
    if water_in_convex_hull:
        if box is not None:
            raise ValueError("the convex-hull option has no minimum-image form")
        nD = pt.nD
        sel = pos[pt.entry_atom[:fw]]
        wpos = pos[pt.water_atoms]
        e0 = np.zeros(0, np.int64)
        e1 = np.zeros(0, np.int64)
        if nD > 0 and pt.nA > 0:
            ia, jd = ball_pairs(pos[pt.acceptors], pos[pt.donors], pt.distance, backend)
            e0, e1 = ia + nD, jd
        true, label = hull_water_labels(pt, pos, max_water, backend)
        tpos = wpos[true]
        si, lj = ball_pairs(sel, tpos, pt.distance, backend)           # positions of the in-hull waters ...
        wi, wj2 = self_pairs(tpos, pt.distance, backend)
        e0 = np.concatenate([e0, label[wi] + fw, si])                  # ... names of the unfiltered list
        e1 = np.concatenate([e1, label[wj2] + fw, label[lj] + fw])
        if pt.check_angle:
            e0, e1 = _angle_filter(pt, e0, e1, np.vstack([sel, wpos]), pos[pt.h_atom], pt.cut_angle, cos_threshold, None)
    else:
        e0, e1 = frame_pairs(pt, pos, 'water_around', around, False, backend, cos_threshold, False, box)
    lo, hi = np.minimum(e0, e1), np.maximum(e0, e1)          # np.sort of every pair (waterwire.py:245-247)
    da = hi < fw
    ww = lo >= fw
    sw = ~da & ~ww
    return (lo[da], hi[da]), (lo[sw], hi[sw]), (lo[ww], hi[ww])


def wires_of_frame(pt, node_of, edges, max_water, allow_direct_bonds=True):
    """waterwire.py:250-309 for one frame. Returns (wires, directs): wires = {(source node, target node): waters in the
    shortest wire}, source < target, in the order the reference first meets them; directs = [(donor entry,
    acceptor entry)] in the reference's row order (acceptor-major, waterwire.py:230,245)."""
    (d_lo, d_hi), (s_e, s_w), (w_a, w_b) = edges
    adj = {}
    for a, b in zip(w_a.tolist(), w_b.tolist()):
        adj.setdefault(a, []).append(b)
        adj.setdefault(b, []).append(a)
    s_n = node_of[s_e]
    by_source = {}
    for n, w in zip(s_n.tolist(), s_w.tolist()):
        by_source.setdefault(n, []).append(w)
    wires = {}
    for source in sorted(by_source):                          # residues = unique(local_hbonds[:, 0]), ascending
        depth = {}
        frontier = []
        for w in by_source[source]:
            if w not in depth and max_water >= 1:
                depth[w] = 1
                frontier.append(w)
        d = 1
        while frontier and d < max_water:                     # single_source_shortest_path(g, source, max_water)
            d += 1
            nxt = []
            for u in frontier:
                for v in adj.get(u, ()):
                    if v not in depth:
                        depth[v] = d
                        nxt.append(v)
            frontier = nxt
        for target, w in zip(s_n.tolist(), s_w.tolist()):     # rows of local_hbonds whose water was reached
            if target <= source or w not in depth:            # already_checked holds every smaller source
                continue
            k = depth[w]                                      # waters in this wire = len(wire) - 2
            key = (source, target)
            if key not in wires or k < wires[key]:
                wires[key] = k
    directs = []
    if allow_direct_bonds:
        order = np.argsort(d_hi, kind='stable')               # rows come acceptor by acceptor
        directs = list(zip(d_lo[order].tolist(), d_hi[order].tolist()))
    return wires, directs


class WireResult:
    def __init__(self, pt, wire_lengths):
        self.topology = pt
        self.wire_lengths = wire_lengths
        self.initial_results = {k: v != np.inf for k, v in wire_lengths.items()}
        self.filtered_results = self.initial_results
        self.nb_frames = pt.nb_frames
        self.initial_graph = dict2graph(self.initial_results, pt.residuewise)
        self.filtered_graph = self.initial_graph

    def compute_average_water_per_wire(self):                 # waterwire.py:359-362
        return {k: v[v < np.inf].mean() for k, v in self.wire_lengths.items()}


def run_wires(universe, selection, max_water=5, allow_direct_bonds=True, backend='kdtree', cos_threshold=None,
              box=None, water_in_convex_hull=False, **ctor):
    """WireAnalysis(...).set_water_wires(max_water, allow_direct_bonds, water_in_convex_hull).
    Returns a WireResult: `wire_lengths` = {key: float64[nb_frames]} (inf = no wire, 0 = direct H-bond, k = k waters)
    with the reference's key orientation (first discovery decides, waterwire.py:277-279, 299-301) and key order.
    The wire paths themselves (`hashs`, `hash_table`: Python `hash(str(path))`, salted per process) are not restated.

    Divergence: frames in which one of the three pair classes is empty make the reference raise
    (waterwire.py:228,231,248); here they are frames that contribute nothing from that class."""
    pt = PortTopology(universe, selection, **ctor)
    node_of = wire_node_of_entry(pt)
    ids = pt.all_ids
    results = {}
    frames = universe.trajectory.frames
    for col, f in enumerate(pt.frame_indices):
        pos = np.asarray(frames.frame(f), dtype=np.float32)
        bk = None if box is None else (np.asarray(box)[col] if np.asarray(box).ndim == 3 else np.asarray(box))
        edges = wire_frame_edges(pt, pos, max_water, backend, cos_threshold, bk, water_in_convex_hull)
        wires, directs = wires_of_frame(pt, node_of, edges, max_water, allow_direct_bonds)
        for (s, t), k in wires.items():
            key = ids[s] + ':' + ids[t]
            rev = ids[t] + ':' + ids[s]
            if rev in results:
                key = rev
            row = results.get(key)
            if row is None:
                row = results[key] = np.full(pt.nb_frames, np.inf)
            if k < row[col]:
                row[col] = k
        for a, b in directs:
            key = ids[a] + ':' + ids[b]
            rev = ids[b] + ':' + ids[a]
            if rev in results:
                key = rev
            row = results.get(key)
            if row is None:
                row = results[key] = np.full(pt.nb_frames, np.inf)
            row[col] = 0
    return WireResult(pt, results)


# ------------------------------------------------------------------------------------------------
# HydrationAnalysis.set_presence_around (hydration.py:71-107), per_residue=False:
# the waters (water group, basic.py:62) within `around_distance` of ANY atom of the selection, per frame.
# ------------------------------------------------------------------------------------------------
def run_presence_around(universe, selection, around_distance, start=None, stop=None, step=1, backend='kdtree',
                        per_residue=False):
    """Returns {water id: bool[nb_frames]} like HydrationAnalysis.initial_results after set_presence_around;
    per_residue=True (hydration.py:88-94): keys 'selection-atom residue id:water id' for every (selected atom, water)
    pair within the distance, the reversed key when that one already exists."""
    topo = universe.topology
    sel = universe.select_atoms(selection).indices
    if len(sel) == 0:
        raise AssertionError('No atoms match the selection')
    waters = universe.select_atoms(WATER_DEFINITION).indices
    water_ids = [_id(topo, i, False) for i in waters]
    sel_ids = [_id(topo, i, False) for i in sel]
    sl = slice(start if isinstance(start, int) else None, stop if isinstance(stop, int) else None, step)
    frame_indices = list(universe.trajectory.frame_indices(sl))
    frames = universe.trajectory.frames
    result = {}
    for col, f in enumerate(frame_indices):
        pos = np.asarray(frames.frame(f), dtype=np.float32)
        si, wj = ball_pairs(pos[sel], pos[waters], float(around_distance), backend)
        if per_residue:
            order = np.lexsort((wj, si))                       # rows come selection atom by selection atom
            keys = []
            for i, j in zip(si[order].tolist(), wj[order].tolist()):
                a, b = sel_ids[i], water_ids[j]
                keys.append(b + ':' + a if (b + ':' + a) in result else a + ':' + b)
        else:
            keys = [water_ids[w] for w in np.unique(wj)]
        for key in keys:
            row = result.get(key)
            if row is None:
                row = result[key] = np.zeros(len(frame_indices), dtype=bool)
            row[col] = True
    return result


# ------------------------------------------------------------------------------------------------
# ConservedGraph.get_conserved_graph, graph_type "hbond" (cgraphs/conservedgraph.py:114-170)
# ------------------------------------------------------------------------------------------------
WATER_TYPES = ("HOH", "TIP3", "TIP4")                          # cgraphs/helperfunctions.py:47


def conserved_graph(graphs, water_positions, reference_coordinates, conservation_threshold=0.9, eps=1.5):
    """graphs: one networkx graph per structure; water_positions: per structure {resid: position} of its waters
    (conservedgraph.py:118-124). Water ends of an edge that lie within eps of a cluster centre 'X-w-<k>' are renamed
    to it (the last matching centre wins, conservedgraph.py:131-149); nodes are counted under their own names. Returns
    (conserved node names, conserved edges as frozensets / 1-tuples for renamed self pairs): present in at least
    round(n * conservation_threshold) structures; water nodes only when some conserved edge names them."""
    nodes, edges = [], []
    centres = [(k, c) for k, c in (reference_coordinates or {}).items() if k.startswith("X-w")]
    for g, waters in zip(graphs, water_positions):
        nodes += list(g.nodes)
        for edge in g.edges:
            edge = [edge[0], edge[1]]
            for i in (0, 1):
                if edge[i].split("-")[1] in WATER_TYPES:
                    n = edge[i].split("-")[2]
                    for key, cc in centres:
                        w = waters.get(int(n))
                        if w is not None and ((w[0] - cc[0]) ** 2 + (w[1] - cc[1]) ** 2 + (w[2] - cc[2]) ** 2 <= float(eps) ** 2):
                            edge[i] = "X-w-" + key.split("-")[-1]
            edges.append(frozenset(edge) if edge[0] != edge[1] else (edge[0],))
    th = np.round(len(graphs) * conservation_threshold)
    un, cn = np.unique(nodes, return_counts=True) if nodes else (np.zeros(0, str), np.zeros(0, int))
    count = {}
    for e in edges:
        count[e] = count.get(e, 0) + 1
    c_edges = {e for e, c in count.items() if c >= th}
    named = {x for e in c_edges for x in e}
    c_nodes = {str(x) for x, c in zip(un, cn) if c >= th and (x.split("-")[1] not in WATER_TYPES or x in named)}
    return c_nodes, c_edges
