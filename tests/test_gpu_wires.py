"""Water wires on the GPU: `cgraphs_b200.mdhbond.WireAnalysis.set_water_wires` against the golden
vectors of the live reference and against the oracle on larger seeded systems - wire lengths per residue pair and
frame, key orientation, graph, averages; fused and general search kernels; capacity overflows and table growth."""
import itertools

import numpy as np
import pytest

from cgraphs_b200 import synth
from tests.helpers import assert_same_wires, graph_edges, load_golden_wires, wire_lists

pytestmark = pytest.mark.gpu


def _gpu(u, sel, max_water, direct=True, options=None, **kw):
    from cgraphs_b200.mdhbond import WireAnalysis
    wa = WireAnalysis(sel, u, native_options=options or {}, **kw)
    wa.set_water_wires(max_water=max_water, allow_direct_bonds=direct)
    return wa


def _port(u, sel, max_water, direct=True, **kw):
    from oracle import hbond_port
    return hbond_port.run_wires(u, sel, max_water=max_water, allow_direct_bonds=direct, **kw)


def _n_runs():
    return len(load_golden_wires()[1])


@pytest.mark.parametrize("i", range(_n_runs()))
def test_wires_equal_golden(i):
    u, runs = load_golden_wires()
    r = runs[i]
    for fused in (1, 0):
        wa = _gpu(u, r["selection"], r["max_water"], r["allow_direct_bonds"], {"fused": fused}, **r["kwargs"])
        assert_same_wires(wire_lists(wa.wire_lengths), r["wire_lengths"], "run %d fused=%d" % (i, fused))
        assert sorted(sorted(e) for e in wa.filtered_graph.edges()) == r["graph_edges"]
        assert wa.nb_frames == r["nb_frames"]


@pytest.mark.parametrize("seed", [0, 1])
def test_wires_equal_oracle_random(seed):
    s = synth.make_system(3000, 36, 100 + seed, hydrogens=True, water="TIP3", n_chains=2)
    u = s.universe(5)
    for sel, mw, ca, direct in itertools.product(["protein", "protein or resname TIP3"], [0, 1, 3, 5], [True, False],
                                                 [True, False]):
        if sel != "protein" and (mw > 3 or not direct):
            continue
        kw = dict(check_angle=ca, additional_donors=['N'], additional_acceptors=['O'], add_donors_without_hydrogen=not ca)
        ref = _port(u, sel, mw, direct, **kw)
        got = _gpu(u, sel, mw, direct, **kw)
        what = "%s max_water=%d angle=%s direct=%s" % (sel, mw, ca, direct)
        assert_same_wires(wire_lists(got.wire_lengths), wire_lists(ref.wire_lengths), what)
        assert graph_edges(got.filtered_graph) == graph_edges(ref.filtered_graph)
        a, b = got.compute_average_water_per_wire(), ref.compute_average_water_per_wire()
        assert a.keys() == b.keys() and all(a[k] == b[k] for k in a)
        assert {k: v.tobytes() for k, v in got.initial_results.items()} == \
               {k: v.tobytes() for k, v in ref.initial_results.items()}


def test_wires_membrane_like_many_frames():
    """C3-like system: the 10.5 A water shell does not fit the fused kernel's shared memory - general kernels,
    several search + BFS rounds, key orientation decided by first discovery across frames."""
    s, cfg = synth.config("C3", scale=0.2)
    u = s.universe(40)
    for ca in (True, False):
        kw = dict(check_angle=ca, add_donors_without_hydrogen=not ca, additional_donors=['N'], additional_acceptors=['O'])
        ref = _port(u, "protein", 5, **kw)
        assert len(ref.wire_lengths) > 20
        got = _gpu(u, "protein", 5, **kw)
        assert_same_wires(wire_lists(got.wire_lengths), wire_lists(ref.wire_lengths), "angle=%s" % ca)


def test_wires_capacity_overflows_are_retried():
    s = synth.make_system(3000, 36, 105, hydrogens=True, water="TIP3", n_chains=2)
    u = s.universe(6)
    kw = dict(check_angle=False, add_donors_without_hydrogen=True, additional_donors=['N'], additional_acceptors=['O'])
    ref = wire_lists(_port(u, "protein", 4, **kw).wire_lengths)
    assert len(ref) > 50
    for options in ({"wire_edge_cap": 256}, {"table_capacity": 16}, {"wire_edge_cap": 300, "table_capacity": 32, "fused": 0},
                    {"fused_lw_cap": 64}, {"wire_smem_words": 0}, {"wire_smem_words": 700}, {"fused": 0, "frame_pairs": 0}):
        got = _gpu(u, "protein", 4, options=options, **kw)
        assert_same_wires(wire_lists(got.wire_lengths), ref, str(options))


def test_wires_then_hbonds_on_one_object():
    """The wire mode of a context is switched off again for the H-bond methods."""
    from oracle import hbond_port
    s = synth.make_system(2000, 24, 31, hydrogens=True, water="TIP3")
    u = s.universe(3)
    kw = dict(additional_donors=['N'], additional_acceptors=['O'])
    wa = _gpu(u, "protein", 3, **kw)
    assert_same_wires(wire_lists(wa.wire_lengths), wire_lists(_port(u, "protein", 3, **kw).wire_lengths))
    wa.set_hbonds_in_selection_and_water_around(3.5)
    ref = hbond_port.run(u, "protein", "water_around", around=3.5, **kw).initial_results
    assert {k: v.tobytes() for k, v in wa.initial_results.items()} == {k: v.tobytes() for k, v in ref.items()}
    wa.set_water_wires(max_water=2)
    assert_same_wires(wire_lists(wa.wire_lengths), wire_lists(_port(u, "protein", 2, **kw).wire_lengths))


def test_wires_unsupported_options_raise():
    from cgraphs_b200.mdhbond import WireAnalysis
    s = synth.make_system(1500, 12, 5, hydrogens=True, water="TIP3")
    u = s.universe(1)
    with pytest.raises(NotImplementedError):
        WireAnalysis("protein or resname TIP3", u).set_water_wires(water_in_convex_hull=True)


def test_wide_shells_general_path_local_water_kernels():
    """Wide around radii on the general path: the local-water test with the selection grid staged in shared memory
    (k_local_water_smem) and from global memory (k_local_water) against the oracle."""
    from cgraphs_b200.mdhbond import HbondAnalysis
    from oracle import hbond_port
    s = synth.make_system(6000, 40, 611, n_chains=2)
    u = s.universe(3)
    kw = dict(additional_donors=['N'], additional_acceptors=['O'])
    for around in (3.5, 8.75, 14.0):
        ref = hbond_port.run(u, "protein", "water_around", around=around, **kw).initial_results
        assert len(ref) > 30
        for lw_smem in (1, 0):
            hba = HbondAnalysis("protein", u, native_options={"fused": 0, "lw_smem": lw_smem}, **kw)
            hba.set_hbonds_in_selection_and_water_around(around)
            assert {k: v.tobytes() for k, v in hba.initial_results.items()} == {k: v.tobytes() for k, v in ref.items()}, around


@pytest.mark.parametrize("block", range(4))
def test_wires_random_sweep(block):
    """Small random systems x random parameters (selection, max_water, criteria, direct bonds, kernel options)."""
    rng = np.random.default_rng(7300 + block)
    for case in range(6):
        n_atoms, n_res = int(rng.integers(700, 3500)), int(rng.integers(4, 30))
        s = synth.make_system(n_atoms, n_res, int(rng.integers(0, 10 ** 6)), n_chains=int(rng.integers(1, 3)),
                              dimensions_kind=["ortho", "tric"][int(rng.integers(2))])
        u = s.universe(int(rng.integers(1, 6)))
        ca = bool(rng.integers(2))
        kw = dict(check_angle=ca, add_donors_without_hydrogen=not ca, distance=float(np.round(rng.uniform(2.8, 4.0), 2)),
                  cut_angle=float(np.round(rng.uniform(30.0, 100.0), 1)))
        if rng.integers(2):
            kw.update(additional_donors=['N'], additional_acceptors=['O'])
        sel = ["protein", "protein or resname TIP3", "segid PA"][int(rng.integers(3))]
        mw, direct = int(rng.integers(0, 7)), bool(rng.integers(2))
        options = {"fused": int(rng.integers(2)), "frame_pairs": int(rng.integers(2)), "lw_smem": int(rng.integers(2))}
        try:
            ref = _port(u, sel, mw, direct, **kw)
        except AssertionError:
            continue                                    # (selection without donors / acceptors)
        got = _gpu(u, sel, mw, direct, options, **kw)
        assert_same_wires(wire_lists(got.wire_lengths), wire_lists(ref.wire_lengths),
                          "block %d case %d: %s max_water=%d %s %s" % (block, case, sel, mw, kw, options))


# ---- convex-hull option: the call cgraphs itself makes (proteingraphanalyser.py:484, water_in_convex_hull=max_water) ------
@pytest.mark.parametrize("i", range(len(load_golden_wires("wires_hull")[1])))
def test_hull_wires_equal_golden(i):
    from cgraphs_b200.mdhbond import WireAnalysis
    u, runs = load_golden_wires("wires_hull")
    r = runs[i]
    for fused in (1, 0):
        wa = WireAnalysis(r["selection"], u, native_options={"fused": fused}, **r["kwargs"])
        wa.set_water_wires(max_water=r["max_water"], allow_direct_bonds=r["allow_direct_bonds"], water_in_convex_hull=r["max_water"])
        assert_same_wires(wire_lists(wa.wire_lengths), r["wire_lengths"], "hull run %d fused=%d" % (i, fused))
        assert sorted(sorted(e) for e in wa.filtered_graph.edges()) == r["graph_edges"]


def test_hull_wires_equal_oracle_membrane_like():
    """C3-like system, several batches, then the plain option on the same object (the virtual water lists are dropped)."""
    from cgraphs_b200.mdhbond import WireAnalysis
    from oracle import hbond_port
    s, cfg = synth.config("C3", scale=0.1)
    u = s.universe(12)
    kw = dict(additional_donors=['N'], additional_acceptors=['O'])
    wa = WireAnalysis("protein", u, native_options={"max_batch_frames": 5}, **kw)
    for mw in (5, 2):
        ref = hbond_port.run_wires(u, "protein", max_water=mw, water_in_convex_hull=True, **kw)
        wa.set_water_wires(max_water=mw, water_in_convex_hull=mw)
        assert_same_wires(wire_lists(wa.wire_lengths), wire_lists(ref.wire_lengths), "hull C3-like max_water=%d" % mw)
        assert len(ref.wire_lengths) > 10
    plain = hbond_port.run_wires(u, "protein", max_water=5, **kw)
    wa.set_water_wires(max_water=5)
    assert_same_wires(wire_lists(wa.wire_lengths), wire_lists(plain.wire_lengths), "plain after hull")


def test_wires_atomwise_keys():
    """residuewise=False (waterwire.py:60-65, 277-301): wires named by the first entry of each residue, direct bonds by
    the atoms themselves; plain and convex-hull option, against the oracle (pinned to the live reference on CPU)."""
    from cgraphs_b200.mdhbond import WireAnalysis
    from oracle import hbond_port
    s = synth.make_system(3000, 40, 5, hydrogens=True, water="TIP3", n_chains=2)
    u = s.universe(5)
    for mw, ca, hull in itertools.product([1, 3], [True, False], [False, True]):
        kw = dict(check_angle=ca, add_donors_without_hydrogen=not ca, residuewise=False,
                  additional_donors=['N'], additional_acceptors=['O'])
        ref = hbond_port.run_wires(u, "protein", max_water=mw, water_in_convex_hull=hull, **kw)
        wa = WireAnalysis("protein", u, **kw)
        wa.set_water_wires(max_water=mw, water_in_convex_hull=hull)
        what = "atomwise max_water=%d angle=%s hull=%s" % (mw, ca, hull)
        assert_same_wires(wire_lists(wa.wire_lengths), wire_lists(ref.wire_lengths), what)
        assert graph_edges(wa.filtered_graph) == graph_edges(ref.filtered_graph), what
        assert any(len(k.split(':')[0].split('-')) == 4 for k in ref.wire_lengths)
