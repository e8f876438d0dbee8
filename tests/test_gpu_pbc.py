"""Minimum-image modes (HB_PBC_ORTHO / HB_PBC_TRICLINIC): an extension the reference does not have.

* On systems whose atoms keep a vacuum skin from every cell face no periodic image can come within the
  cut-off, so the minimum-image search must return exactly what the reference's non-periodic search returns
   - checked against the oracle's reference semantics.
* On systems that do cross the faces the kernels are checked bit for bit against the oracle's restatement of
  the same minimum-image definition (oracle/hbond_port.py), fused and general kernels alike.
"""
import itertools

import numpy as np
import pytest

from cgraphs_b200 import synth
from cgraphs_b200.universe import Universe
from tests.helpers import assert_same_results

pytestmark = pytest.mark.gpu


def _run(u, selection, method, options=None, around=3.5, **kw):
    from cgraphs_b200.mdhbond import HbondAnalysis
    hba = HbondAnalysis(selection, u, native_options=options or {}, **kw)
    if method == "in_selection":
        hba.set_hbonds_in_selection()
    else:
        hba.set_hbonds_in_selection_and_water_around(around)
    return hba


def _port(u, selection, method, around=3.5, **kw):
    from oracle import hbond_port
    return hbond_port.run(u, selection, method, around=around, **kw)


def _periodic_copy(s, n_frames, offset_frac=(0.47, 0.52, 0.49)):
    """The system translated so that the protein straddles the cell faces, every residue (molecule) wrapped
    into the cell as a whole; returns (Universe, float32 lattice rows)."""
    from oracle import hbond_port
    box = hbond_port.box_rows(s.dimensions)
    h = box.astype(np.float64)
    xyz = s.frames(0, n_frames).astype(np.float64) + np.asarray(offset_frac) @ h
    t = s.topology
    res_key = np.array(["%s|%s|%s" % (a, b, c) for a, b, c in zip(t.segids, t.resids, t.resnames)])
    _, res_idx = np.unique(res_key, return_inverse=True)
    n_res = res_idx.max() + 1
    counts = np.bincount(res_idx, minlength=n_res)[:, None]
    hinv = np.linalg.inv(h)
    out = np.empty_like(xyz)
    for f in range(n_frames):
        cen = np.stack([np.bincount(res_idx, weights=xyz[f][:, k], minlength=n_res) for k in range(3)], axis=1) / counts
        shift = np.floor(cen @ hinv) @ h
        out[f] = xyz[f] - shift[res_idx]
    return Universe(t, out.astype(np.float32), s.dimensions), box


@pytest.mark.parametrize("cfg_name,scale", [("C2", 0.3), ("C3", 0.3)])
def test_pbc_equals_reference_semantics_on_skin_systems(cfg_name, scale):
    s, cfg = synth.config(cfg_name, scale=scale)
    u = s.universe(4)
    ref = _port(u, "protein", "water_around").initial_results          # reference semantics: no periodic images
    assert len(ref) > 10
    for fused in (1, 0):
        got = _run(u, "protein", "water_around", {"fused": fused}, pbc=True)
        assert_same_results(got.initial_results, ref, "%s fused=%d" % (cfg_name, fused))
        if fused:
            assert got.last_run_stats["fused_frames"] == 4


@pytest.mark.parametrize("kind", ["ortho", "tric"])
def test_pbc_matches_oracle_across_cell_faces(kind):
    s = synth.make_system(6000, 40, 611, dimensions_kind=kind, n_chains=2)
    u, box = _periodic_copy(s, 3)
    pos = u.trajectory.frames.block(0, 3)
    frac = pos.astype(np.float64) @ np.linalg.inv(box.astype(np.float64))
    assert frac.min() < 0.02 and frac.max() > 0.98                    # atoms really sit at the faces
    for sel, method, ca in itertools.product(["protein", "protein or resname TIP3"], ["in_selection", "water_around"],
                                             [True, False]):
        kw = dict(check_angle=ca, additional_donors=['N'], additional_acceptors=['O'])
        ref = _port(u, sel, method, box=box, **kw).initial_results
        nopbc = _port(u, sel, method, **kw).initial_results
        assert len(ref) > len(nopbc)                                   # the faces do cut bonds without images
        for fused in (1, 0):
            got = _run(u, sel, method, {"fused": fused}, pbc=True, **kw)
            assert_same_results(got.initial_results, ref, "%s %s %s %s fused=%d" % (kind, sel, method, ca, fused))


def test_pbc_small_cell_few_cells_per_axis():
    """Cut-offs close to half the cell: one or two cells per axis, every neighbour list wraps onto itself."""
    s = synth.make_system(1500, 6, 77, dimensions_kind="tric")
    u, box = _periodic_copy(s, 2)
    edge = float(min(box[0, 0], box[1, 1], box[2, 2]))
    for dist, around in [(3.5, 3.5), (0.3 * edge, 0.45 * edge), (0.45 * edge, 0.3 * edge)]:
        kw = dict(distance=dist, check_angle=True)
        for sel in ("protein", "protein or resname TIP3"):
            ref = _port(u, sel, "water_around", around=around, box=box, **kw).initial_results
            for fused in (1, 0):
                got = _run(u, sel, "water_around", {"fused": fused}, around=around, pbc=True, **kw)
                assert_same_results(got.initial_results, ref, "d=%.2f a=%.2f %s fused=%d" % (dist, around, sel, fused))


def test_pbc_argument_errors():
    from cgraphs_b200.mdhbond import _native
    s = synth.make_system(1500, 6, 78, dimensions_kind="ortho")
    u = s.universe(1)
    with pytest.raises(_native.HbError, match="half"):
        _run(u, "protein", "water_around", pbc=True, distance=0.6 * float(s.dimensions[0]))
    with pytest.raises(_native.HbError, match="orthorhombic"):
        s2 = synth.make_system(1500, 6, 78, dimensions_kind="tric")
        _run(s2.universe(1), "protein", "water_around", pbc="ortho")
    u_nobox = Universe(s.topology, s.frames(0, 1))
    with pytest.raises(ValueError, match="dimensions"):
        _run(u_nobox, "protein", "water_around", pbc=True)
    # the C ABI refuses lattices that are not lower-triangular rows
    hba = _run(u, "protein", "water_around", pbc=True)
    ctx = hba._context()
    bad = np.array([[[30, 0, 1], [0, 30, 0], [0, 0, 30]]], dtype=np.float32)
    ctx.begin(1)
    with pytest.raises(_native.HbError, match="lower-triangular"):
        ctx.submit(s.frames(0, 1), box=bad)


def _ckdtree_periodic_keys(pt, pos, L, method, around):
    """One frame's bond keys with scipy's own periodic cKDTree (boxsize=...) doing every distance query - an independent
    check of the orthorhombic minimum-image search: wrapped float32 coordinates, no angle criterion."""
    from scipy.spatial import cKDTree
    from oracle import hbond_port
    Ld = np.asarray(L, dtype=np.float64)
    wrap = lambda p: np.mod(p.astype(np.float64), Ld)              # cKDTree wants [0, L)
    nD, fw = pt.nD, pt.first_water
    sel = wrap(pos[pt.entry_atom[:fw]])
    ta, td = cKDTree(wrap(pos[pt.acceptors]), boxsize=Ld), cKDTree(wrap(pos[pt.donors]), boxsize=Ld)
    e0, e1 = [], []
    for i, js in enumerate(ta.query_ball_tree(td, pt.distance)):
        for j in js:
            e0.append(i + nD)
            e1.append(j)
    e0, e1 = np.asarray(e0, np.int64), np.asarray(e1, np.int64)
    bb = hbond_port._backbone_flags(pt)
    keep = ~(bb[e0] & bb[e1])
    e0, e1 = e0[keep], e1[keep]
    if method == "water_around":
        wpos = wrap(pos[pt.water_atoms])
        tw, ts = cKDTree(wpos, boxsize=Ld), cKDTree(sel, boxsize=Ld)
        local = np.unique([j for js in ts.query_ball_tree(tw, around) for j in js]).astype(np.int64)
        tl = cKDTree(wpos[local], boxsize=Ld)
        s0, s1 = [], []
        for i, js in enumerate(ts.query_ball_tree(tl, pt.distance)):
            for j in js:
                s0.append(i)
                s1.append(local[j] + fw)
        ww = np.array(sorted(tl.query_pairs(pt.distance)), dtype=np.int64).reshape(-1, 2)
        e0 = np.concatenate([e0, local[ww[:, 0]] + fw, np.asarray(s0, np.int64)])
        e1 = np.concatenate([e1, local[ww[:, 1]] + fw, np.asarray(s1, np.int64)])
    return hbond_port._frame_keys(pt, e0, e1, hbond_port.key_ids(pt, method))


@pytest.mark.parametrize("method", ["in_selection", "water_around"])
def test_ortho_pbc_equals_scipy_periodic_ckdtree(method):
    """Every atom wrapped into the cell on its own (molecules cut by the faces), the protein straddling three faces:
    the per-frame bond sets of the minimum-image kernels equal what scipy's cKDTree(boxsize=...) finds."""
    from oracle import hbond_port
    s, cfg = synth.config("C2", scale=0.3)
    box = hbond_port.box_rows(s.dimensions)
    L = np.array([box[0, 0], box[1, 1], box[2, 2]], dtype=np.float32)
    nf = 4
    xyz = s.frames(0, nf).astype(np.float64) + np.array([0.48, 0.51, 0.5]) * L.astype(np.float64)
    xyz = np.mod(xyz, L.astype(np.float64)).astype(np.float32)
    xyz = np.where(xyz >= L, xyz - L, xyz)                            # (float32 rounding can land exactly on L)
    u = Universe(s.topology, xyz, s.dimensions)
    kw = dict(check_angle=False, add_donors_without_hydrogen=True, additional_donors=['N'], additional_acceptors=['O'])
    pt = hbond_port.PortTopology(u, "protein", **kw)
    want = {}
    for f in range(nf):
        for key in _ckdtree_periodic_keys(pt, xyz[f], L, method, 4.0):
            want.setdefault(key, np.zeros(nf, dtype=bool))[f] = True
    assert len(want) > 8
    for opts in ({"fused": 1}, {"fused": 0}):
        got = _run(u, "protein", method, opts, around=4.0, pbc="ortho", **kw)
        assert_same_results(got.initial_results, want, "scipy periodic cKDTree, %s %s" % (method, opts))
    # and the pairs really cross the faces: without images the result is a different one
    plain = _run(u, "protein", method, around=4.0, **kw)
    assert {k: v.tobytes() for k, v in plain.initial_results.items()} != {k: v.tobytes() for k, v in want.items()}
