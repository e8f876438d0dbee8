"""Parity of the CUDA path (through the C ABI / drop-in class) against the golden vectors of the live
reference and against the CPU oracle on seeded systems.  Bit-exact: same key set, identical bool vectors."""
import itertools

import numpy as np
import pytest

from cgraphs_b200 import synth
from cgraphs_b200.universe import Universe
from tests.helpers import assert_same_results, golden_cases, graph_edges, lists_to_results, load_golden

pytestmark = pytest.mark.gpu

_CACHE = {}


def _fixture(name):
    if name not in _CACHE:
        _CACHE[name] = load_golden(name)
    return _CACHE[name]


def _run_gpu(u, selection, method, around=3.5, excl_bb=True, cos_threshold=None, **kw):
    from cgraphs_b200.mdhbond import HbondAnalysis
    hba = HbondAnalysis(selection, u, cos_threshold=cos_threshold, **kw)
    if method == "in_selection":
        hba.set_hbonds_in_selection(exclude_backbone_backbone=excl_bb)
    else:
        hba.set_hbonds_in_selection_and_water_around(around, exclude_backbone_backbone=excl_bb)
    return hba


def _run_port(u, selection, method, around=3.5, excl_bb=True, **kw):
    from oracle import hbond_port
    return hbond_port.run(u, selection, method, around=around, exclude_backbone_backbone=excl_bb, **kw)


@pytest.mark.parametrize("name,idx", golden_cases())
def test_gpu_matches_golden(name, idx):
    u, runs = _fixture(name)
    run = runs[idx]
    expect = lists_to_results(run["results"], run["nb_frames"])
    hba = _run_gpu(u, run["selection"], run["method"], run["around"], run["excl_bb"],
                   cos_threshold=run["cos_threshold"] if run["kwargs"].get("check_angle", True) else None,
                   **run["kwargs"])
    assert hba.nb_frames == run["nb_frames"]
    assert_same_results(hba.initial_results, expect, "%s[%d]" % (name, idx))
    assert hba.initial_results is hba.filtered_results
    assert graph_edges(hba.filtered_graph) == {frozenset(e) for e in run["graph_edges"]}
    assert all(v.dtype == bool and v.shape == (run["nb_frames"],) for v in hba.initial_results.values())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gpu_matches_port_random(seed):
    s = synth.make_system(4000, 40, 300 + seed, hydrogens=True, water="TIP3", n_chains=2)
    u = s.universe(5)
    for sel, method, ca, rw in itertools.product(["protein", "protein or resname TIP3"],
                                                 ["in_selection", "water_around"], [True, False], [True, False]):
        kw = dict(check_angle=ca, residuewise=rw, additional_donors=['N'], additional_acceptors=['O'])
        a = _run_gpu(u, sel, method, **kw)
        b = _run_port(u, sel, method, **kw)
        assert_same_results(a.initial_results, b.initial_results, "%s %s %s %s" % (sel, method, ca, rw))
        assert graph_edges(a.filtered_graph) == graph_edges(b.filtered_graph)


@pytest.mark.parametrize("cfg_name,scale,frames", [("C1", 1.0, 1), ("C2", 1.0, 8), ("C3", 0.5, 4), ("C3", 1.0, 2),
                                                   ("C4", 1.0, 2)])
def test_gpu_matches_port_baseline_configs(cfg_name, scale, frames):
    s, cfg = synth.config(cfg_name, scale=scale)
    u = s.universe(frames, static=cfg["static"])
    kw = dict(check_angle=cfg["check_angle"], add_donors_without_hydrogen=cfg["add_donors_without_hydrogen"])
    for method in ("in_selection", "water_around"):
        a = _run_gpu(u, cfg["selection"], method, cfg["around"], **kw)
        b = _run_port(u, cfg["selection"], method, cfg["around"], **kw)
        assert len(b.initial_results) > 0
        assert_same_results(a.initial_results, b.initial_results, "%s %s" % (cfg_name, method))


def test_gpu_all_waters_in_selection_c2():
    s, cfg = synth.config("C2", scale=0.4)
    u = s.universe(3)
    a = _run_gpu(u, "protein or resname TIP3", "in_selection")
    b = _run_port(u, "protein or resname TIP3", "in_selection")
    assert len(b.initial_results) > 1000
    assert_same_results(a.initial_results, b.initial_results)


@pytest.mark.parametrize("seed", [5, 6])
def test_gpu_kat_other_seeds(seed):
    """Fresh adversarial systems (not the committed ones) against the oracle on this host."""
    from cgraphs_b200.mdhbond.calib import cos_threshold
    s = synth.kat_distance_system(1200, seed)
    u = s.universe(1, static=True)
    kw = dict(check_angle=False, add_donors_without_hydrogen=True)
    a, b = _run_gpu(u, "resname TIP3", "in_selection", **kw), _run_port(u, "resname TIP3", "in_selection", **kw)
    assert 300 < len(b.initial_results) < 900
    assert_same_results(a.initial_results, b.initial_results, "kat_distance")
    for cut in (60., 45., 30.):
        s = synth.kat_angle_system(700, seed, cos_threshold=cos_threshold(cut))
        u = s.universe(1, static=True)
        a = _run_gpu(u, "resname TIP3", "water_around", cut_angle=cut)
        b = _run_port(u, "resname TIP3", "water_around", cut_angle=cut)
        assert 200 < len(b.initial_results) < 600
        assert_same_results(a.initial_results, b.initial_results, "kat_angle %s" % cut)


def test_gpu_invariances():
    """Translation by a float32-exact offset, atom-order-preserving frame permutation, frame slicing."""
    s = synth.make_system(3000, 30, 77)
    xyz = s.frames(0, 6)
    u = Universe(s.topology, xyz)
    base = _run_gpu(u, "protein", "water_around").initial_results
    # permuting frames permutes result columns
    perm = np.array([3, 0, 5, 1, 4, 2])
    up = Universe(s.topology, xyz[perm])
    rp = _run_gpu(up, "protein", "water_around").initial_results
    assert set(rp) == set(base)
    for k in base:
        assert np.array_equal(rp[k], base[k][perm])
    # start/stop/step
    sl = _run_gpu(u, "protein", "water_around", start=1, stop=6, step=2).initial_results
    for k, v in sl.items():
        assert np.array_equal(v, base[k][1:6:2])
    assert {k for k, v in base.items() if v[1:6:2].any()} == set(sl)


def test_gpu_empty_and_degenerate_frames():
    """Frames without any candidate pair are all-False columns (documented divergence)."""
    s = synth.make_system(2000, 20, 9)
    xyz = s.frames(0, 3)
    # frame 1: blow the system up so nothing is within 3.5 A (hydrogens stay bonded in frame 0)
    centre = xyz[1].mean(axis=0)
    xyz[1] = (xyz[1] - centre) * 40.0 + centre
    u = Universe(s.topology, xyz)
    a = _run_gpu(u, "protein", "water_around").initial_results
    b = _run_port(u, "protein", "water_around").initial_results
    assert_same_results(a, b)
    assert not any(v[1] for v in a.values())
    # identical coordinates for everything: all pairs at distance 0
    xyz0 = np.zeros_like(xyz[:1])
    u0 = Universe(s.topology, np.concatenate([xyz[:1], xyz0]))
    a = _run_gpu(u0, "protein", "in_selection", check_angle=False).initial_results
    b = _run_port(u0, "protein", "in_selection", check_angle=False).initial_results
    assert_same_results(a, b)


def test_gpu_error_behaviour():
    from cgraphs_b200.mdhbond import HbondAnalysis
    s = synth.make_system(2500, 40, 3)
    u = s.universe(1)
    with pytest.raises(AssertionError, match="No selection string"):
        HbondAnalysis(None, u)
    with pytest.raises(AssertionError, match="No structure file path"):
        HbondAnalysis("protein", None)
    with pytest.raises(AssertionError, match="No atoms match the selection"):
        HbondAnalysis("resname XXX", u)
    with pytest.raises(AssertionError, match="neither donors nor acceptors"):
        HbondAnalysis("name CA", u)
    # carboxylate / carbonyl oxygens carry no hydrogen: acceptors only
    names = set(s.topology.names)
    acc_only = [n for n in ("OD1", "OE1", "OD2", "OE2") if n in names]
    h = HbondAnalysis("name " + " ".join(acc_only), u)
    assert h._nb_donors == 0 and h._nb_acceptors > 0
    with pytest.raises(AssertionError, match="No donors in the selection"):
        h.set_hbonds_in_selection()


def test_gpu_table_growth_and_batches():
    """Tiny initial bond table and tiny batches: the table must grow and batches re-run without loss."""
    from cgraphs_b200.mdhbond import HbondAnalysis
    s, cfg = synth.config("C2", scale=0.3)
    u = s.universe(70)
    ref = _run_port(u, "protein or resname TIP3", "water_around").initial_results
    hba = HbondAnalysis("protein or resname TIP3", u, batch_bytes=8 << 20)
    ctx = hba._context()
    ctx.set_option("table_capacity", 1024)
    ctx.set_option("max_batch_frames", 7)
    hba.set_hbonds_in_selection_and_water_around(3.5)
    assert_same_results(hba.initial_results, ref)
    assert hba.last_run_stats["table_grows"] >= 1
    assert hba.last_run_stats["frames"] == 70


def test_gpu_device_resident_and_pinned_paths():
    """hb_submit_frames from pinned memory and hb_submit_frames_device give the same matrix as pageable."""
    import torch
    from cgraphs_b200.mdhbond import HbondAnalysis, _native
    from cgraphs_b200.mdhbond.topology import pack_systems
    s = synth.make_system(6000, 50, 31)
    nf = 40
    xyz = s.frames(0, nf)
    u = Universe(s.topology, xyz)
    hba = HbondAnalysis("protein", u)
    hba.set_hbonds_in_selection_and_water_around(3.5)
    ctx = hba._context()
    # same context, same topology/params: re-run from other sources
    outs = []
    ctx.begin(nf)
    ctx.submit(xyz)
    outs.append(ctx.results())
    pin = _native.PinnedBuffer(xyz.shape)
    pin.array[...] = xyz
    ctx.begin(nf)
    ctx.submit_raw(pin.ptr, xyz.shape[1], nf)
    outs.append(ctx.results())
    d = torch.from_numpy(xyz).cuda()
    torch.cuda.synchronize()
    ctx.begin(nf)
    ctx.submit_device(d.data_ptr(), xyz.shape[1], nf // 2)
    ctx.submit_device(d.data_ptr() + (nf // 2) * xyz.shape[1] * 12, xyz.shape[1], nf - nf // 2)
    outs.append(ctx.results())
    pin.free()
    for k, w in outs[1:]:
        assert np.array_equal(k, outs[0][0]) and np.array_equal(w, outs[0][1])
    assert len(outs[0][0]) == len(hba.initial_results)
    counts = ctx.counts(len(outs[0][0]))
    bits = np.unpackbits(outs[0][1].view(np.uint8), axis=1, bitorder="little")[:, :nf]
    assert np.array_equal(counts, bits.sum(axis=1))


def test_gpu_structure_batch_c5():
    """Config C5 in miniature: independent static structures searched in one batched launch."""
    from cgraphs_b200.mdhbond import HbondBatch
    systems = synth.c5_structures(12, seed0=5000)
    unis = [s.universe(1, static=True) for s in systems]
    kw = dict(check_angle=False, add_donors_without_hydrogen=True)
    batch = HbondBatch(unis, "protein", **kw)
    for method in ("in_selection", "water_around"):
        if method == "in_selection":
            res = batch.set_hbonds_in_selection()
        else:
            res = batch.set_hbonds_in_selection_and_water_around(3.5)
        for u, r in zip(unis, res):
            b = _run_port(u, "protein", method, **kw)
            assert_same_results(r.initial_results, b.initial_results, "batch %s" % method)
            assert graph_edges(r.filtered_graph) == graph_edges(b.filtered_graph)
    assert batch.occupancy.shape[1] == 12


def test_gpu_dump_and_restore(tmp_path):
    from cgraphs_b200.mdhbond import HbondAnalysis
    s = synth.make_system(2000, 20, 13)
    u = s.universe(3)
    hba = HbondAnalysis("protein", u)
    hba.set_hbonds_in_selection_and_water_around(3.5)
    f = str(tmp_path / "hba.pkl")
    hba.dump_to_file(f)
    back = HbondAnalysis(restore_filename=f)
    assert_same_results(back.initial_results, hba.initial_results)
    assert graph_edges(back.filtered_graph) == graph_edges(hba.filtered_graph)
    hba.filter_occupancy(0.5)
    assert all(v.mean() > 0.5 for v in hba.filtered_results.values())
    assert hba.applied_filters['occupancy'] == 0.5
    dup = hba.duplicate()
    assert dup.filtered_results is hba.filtered_results


def test_gpu_occupancy_postprocessing_matches_numpy():
    """filter_occupancy / compute_joint_occupancy / _compute_graph_in_frame on the packed device matrix give what
    the reference's numpy code gives on the bool vectors (network.py:99-107, 369-376, 1186-1198)."""
    import networkx as nx
    from cgraphs_b200.mdhbond import HbondAnalysis
    # odd and even numbers of 32-frame words, a partial tail word
    for n_frames in (5, 45):
        _check_occupancy_postprocessing(n_frames)


def _check_occupancy_postprocessing(n_frames):
    import networkx as nx
    from cgraphs_b200.mdhbond import HbondAnalysis
    s = synth.make_system(4000, 40, 71)
    u = s.universe(n_frames)
    hba = HbondAnalysis("protein", u)
    hba.set_hbonds_in_selection_and_water_around(3.5)
    res = hba.initial_results
    assert hba._device_rows(res) is not None
    # reference formulas on the bool vectors
    occ = {k: float(np.mean(v)) for k, v in res.items()}
    got = hba.occupancies()
    assert got.keys() == occ.keys() and all(got[k] == occ[k] for k in occ)
    thr = float(np.median(list(occ.values())))
    hba.filter_occupancy(thr)
    assert set(hba.filtered_results) == {k for k, o in occ.items() if o > thr}
    assert hba.applied_filters['occupancy'] == thr
    # joint occupancy of the filtered subset (row mask) and of everything
    combined = np.ones(hba.nb_frames, dtype=bool)
    for k in hba.filtered_results:
        combined &= res[k]
    assert hba.compute_joint_occupancy() == combined.mean()
    assert np.array_equal(hba.joint_occupancy_series, combined)
    assert np.array_equal(hba.joint_occupancy_frames, np.nonzero(combined)[0])
    allc = np.ones(hba.nb_frames, dtype=bool)
    for k in res:
        allc &= res[k]
    assert hba.compute_joint_occupancy(use_filtered=False) == allc.mean()
    # per-frame graphs
    for frame in (0, n_frames // 2, n_frames - 1):
        g = hba._compute_graph_in_frame(frame, use_filtered=False)
        ref = nx.Graph()
        ref.add_edges_from((k.split(':')[0], k.split(':')[1]) for k, v in res.items() if v[frame])
        assert graph_edges(g) == graph_edges(ref)
    # a new run invalidates the device rows of the old one: host fallback, same answers
    other = HbondAnalysis("protein", u)
    other._ctx = hba._ctx
    other.set_hbonds_in_selection()
    assert hba._device_rows(res) is None
    assert hba.compute_joint_occupancy(use_filtered=False) == allc.mean()


def test_gpu_batch_conserved_graph():
    """HbondBatch.get_conserved_graph == the counting rule of ConservedGraph.get_conserved_graph
    (conservedgraph.py:114-170, hbond branch without water relabelling) applied to per-structure oracle graphs."""
    from cgraphs_b200.mdhbond import HbondBatch
    # the same protein in several conformations: static structure + frames of its trajectory as separate structures
    s = synth.make_system(3000, 60, 901, hydrogens=False, water="HOH")
    xyz = s.frames(0, 9)
    unis = [Universe(s.topology, xyz[i:i + 1]) for i in range(9)]
    kw = dict(check_angle=False, add_donors_without_hydrogen=True)
    batch = HbondBatch(unis, "protein", **kw)
    batch.set_hbonds_in_selection_and_water_around(3.5)
    for cth in (0.9, 0.5, 0.2):
        nodes, edges = batch.get_conserved_graph(cth)
        # reference rule on per-structure graphs
        all_nodes, all_edges = [], []
        for u in unis:
            g = _run_port(u, "protein", "water_around", **kw).filtered_graph
            all_nodes += list(g.nodes)
            all_edges += [frozenset(e) for e in g.edges]
        th = np.round(len(unis) * cth)
        un, cn = np.unique(all_nodes, return_counts=True)
        ce = {e for e in set(all_edges) if all_edges.count(e) >= th}
        cnodes = {x for x, c in zip(un, cn) if c >= th and (x.split('-')[1] not in ("HOH", "TIP3", "TIP4") or any(x in e for e in ce))}
        assert {frozenset(e) for e in edges} == ce
        assert set(nodes) == cnodes
        assert len(ce) > 0


@pytest.mark.parametrize("residuewise", [True, False])
def test_gpu_ion_contacts(residuewise):
    """ions=[...]: set_add_ion_interactions runs after both set_hbonds_* methods (hbond.py:57-94,135,289)."""
    from cgraphs_b200.mdhbond import HbondAnalysis
    from tests.helpers import ion_system, port_with_ions
    u = ion_system(seed=4, n_frames=5)
    for method, pbc in (("in_selection", False), ("water_around", False), ("water_around", True)):
        ref = port_with_ions(u, "protein", method, ['SOD', 'CLA'], residuewise=residuewise)
        hba = HbondAnalysis("protein", u, ions=['SOD', 'CLA'], residuewise=residuewise, pbc=pbc)
        if method == "in_selection":
            hba.set_hbonds_in_selection()
        else:
            hba.set_hbonds_in_selection_and_water_around(3.5)
        assert_same_results(hba.initial_results, ref, "ions %s pbc=%s" % (method, pbc))
        assert sum(1 for k in ref if k.split(':')[1].split('-')[1] in ('SOD', 'CLA')) > 5
    bare = HbondAnalysis("protein", u)
    with pytest.raises(AssertionError, match="No ions"):
        bare.set_add_ion_interactions()


@pytest.mark.gpu
def test_gpu_ions_behind_an_inert_atom_range():
    """Ions that lie after a long run of atoms no kernel reads (CHARMM order: ... waters, lipids, SOD / CLA): the host
    frames reach the device as atom ranges, and hb_set_ions has to add the ions to them (both submit paths)."""
    from cgraphs_b200.mdhbond import HbondAnalysis
    from cgraphs_b200.universe import Topology
    from tests.helpers import port_with_ions
    s, cfg = synth.config("C3", scale=0.3)
    t = s.topology
    resn = np.asarray(t.resnames)
    fill = np.nonzero(resn == "POPC")[0]
    rest = np.nonzero(resn != "POPC")[0]
    n_ions = 24
    order = np.concatenate([rest, fill])                    # protein, waters, membrane (inert), ions = last filler atoms
    names = [t.names[i] for i in order]
    resnames = [t.resnames[i] for i in order]
    resids = [int(t.resids[i]) for i in order]
    segids = [t.segids[i] for i in order]
    n = len(order)
    for k in range(n_ions):
        i = n - n_ions + k
        names[i] = resnames[i] = "SOD" if k % 2 == 0 else "CLA"
        resids[i] = 7000 + k
        segids[i] = "ION"
    nf = 6
    xyz = s.frames(0, nf)[:, order, :]
    rng = np.random.default_rng(11)
    targets = rng.choice(len(rest), size=n_ions, replace=False)
    for k in range(n_ions):
        xyz[:, n - n_ions + k, :] = xyz[:, targets[k], :] + rng.normal(0, 1.4, size=(nf, 3)).astype(np.float32)
    assert len(fill) - n_ions > 4096 and len(rest) < 0.9 * n
    u = Universe(Topology(names, resnames, resids, segids), np.ascontiguousarray(xyz), s.dimensions)
    ref = port_with_ions(u, "protein", "water_around", ['SOD', 'CLA'])
    assert sum(1 for k in ref if k.split(':')[1].split('-')[1] in ('SOD', 'CLA')) > 5
    for opts in ({}, {"copy_ranges": 0}):
        hba = HbondAnalysis("protein", u, ions=['SOD', 'CLA'], native_options=opts)
        hba.set_hbonds_in_selection_and_water_around(3.5)
        assert_same_results(hba.initial_results, ref, "ions behind a hole %s" % opts)
        if not opts:
            assert hba.last_run_stats["h2d_bytes"] < 0.9 * xyz.nbytes       # ranges were in use
    # planar (DCD record) submission of the same frames
    from cgraphs_b200.mdhbond import _native
    from cgraphs_b200.mdhbond.topology import pack_systems
    hba = HbondAnalysis("protein", u, ions=['SOD', 'CLA'])
    hba.set_hbonds_in_selection_and_water_around(3.5)
    want = dict(hba.initial_results)
    planar = np.ascontiguousarray(xyz.transpose(0, 2, 1))   # [frame][x|y|z][atom]
    ctx = hba._context()
    ctx.begin(nf)
    ctx.submit_planar(planar, n * 12, 0, n * 4, n * 8, n, nf)
    keys, words = ctx.results()
    ctx.begin(nf)
    ctx.submit(np.ascontiguousarray(xyz))
    keys2, words2 = ctx.results()
    assert np.array_equal(keys, keys2) and np.array_equal(words, words2) and len(keys) > 0
    assert want
