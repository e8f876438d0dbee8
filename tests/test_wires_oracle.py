"""Water wires: the CPU restatement `oracle.hbond_port.run_wires` against the golden vectors the
live reference produced (tests/golden/make_golden_wires.py) and against the live reference itself."""
import itertools

import pytest

from cgraphs_b200 import synth
from cgraphs_b200.mdhbond.calib import cos_threshold
from oracle import hbond_port, ref_harness
from tests.helpers import assert_same_wires, graph_edges, load_golden_wires, wire_lists


def _n_runs():
    return len(load_golden_wires()[1])


@pytest.mark.parametrize("i", range(_n_runs()))
def test_port_wires_equal_golden(i):
    u, runs = load_golden_wires()
    r = runs[i]
    for backend, cthr in (("kdtree", None), ("brute", cos_threshold(float(r["kwargs"].get("cut_angle", 60.))))):
        p = hbond_port.run_wires(u, r["selection"], max_water=r["max_water"], allow_direct_bonds=r["allow_direct_bonds"],
                                 backend=backend, cos_threshold=cthr, **r["kwargs"])
        assert_same_wires(wire_lists(p.wire_lengths), r["wire_lengths"], "%d %s" % (i, backend))
        assert list(p.wire_lengths) == list(r["wire_lengths"])          # the reference's dict order too
        assert sorted(sorted(e) for e in p.filtered_graph.edges()) == r["graph_edges"]


@pytest.mark.reference
@pytest.mark.parametrize("seed", [0])
def test_port_wires_equal_reference_random(seed):
    ref = ref_harness.load_reference()
    s = synth.make_system(3000, 36, 100 + seed, hydrogens=True, water="TIP3", n_chains=2)
    u = s.universe(3)
    for sel, mw, ca, direct in itertools.product(["protein", "protein or resname TIP3"], [1, 3, 5], [True, False],
                                                 [True, False]):
        if sel != "protein" and (mw > 3 or not direct):
            continue
        kw = dict(check_angle=ca, additional_donors=['N'], additional_acceptors=['O'], add_donors_without_hydrogen=not ca)
        wa = ref.WireAnalysis(sel, u, **kw)
        wa.set_water_wires(max_water=mw, allow_direct_bonds=direct)
        p = hbond_port.run_wires(u, sel, max_water=mw, allow_direct_bonds=direct, **kw)
        what = "%s %d %s %s" % (sel, mw, ca, direct)
        assert_same_wires(wire_lists(p.wire_lengths), wire_lists(wa.wire_lengths), what)
        assert list(p.wire_lengths) == list(wa.wire_lengths), what
        assert graph_edges(p.filtered_graph) == graph_edges(wa.filtered_graph)
        a, b = p.compute_average_water_per_wire(), wa.compute_average_water_per_wire()
        assert a.keys() == b.keys() and all(a[k] == b[k] for k in a)


@pytest.mark.reference
def test_port_wires_membrane_like():
    ref = ref_harness.load_reference()
    s, cfg = synth.config("C3", scale=0.1)
    u = s.universe(2)
    for ca in (True, False):
        kw = dict(check_angle=ca, add_donors_without_hydrogen=not ca, additional_donors=['N'], additional_acceptors=['O'])
        wa = ref.WireAnalysis("protein", u, **kw)
        wa.set_water_wires(max_water=5)
        p = hbond_port.run_wires(u, "protein", max_water=5, **kw)
        assert len(wa.wire_lengths) > 3
        assert_same_wires(wire_lists(p.wire_lengths), wire_lists(wa.wire_lengths))


# ---- convex-hull option (waterwire.py:218-228; what cgraphs itself always sets, proteingraphanalyser.py:484) ----------
@pytest.mark.parametrize("i", range(len(load_golden_wires("wires_hull")[1])))
def test_port_hull_wires_equal_golden(i):
    u, runs = load_golden_wires("wires_hull")
    r = runs[i]
    p = hbond_port.run_wires(u, r["selection"], max_water=r["max_water"], allow_direct_bonds=r["allow_direct_bonds"],
                             water_in_convex_hull=True, **r["kwargs"])
    assert_same_wires(wire_lists(p.wire_lengths), r["wire_lengths"], "hull %d" % i)
    assert list(p.wire_lengths) == list(r["wire_lengths"])
    assert sorted(sorted(e) for e in p.filtered_graph.edges()) == r["graph_edges"]


@pytest.mark.reference
def test_port_hull_wires_equal_reference_random():
    import warnings
    warnings.filterwarnings("ignore")
    ref = ref_harness.load_reference()
    s = synth.make_system(3000, 40, 5, hydrogens=True, water="TIP3", n_chains=2)
    u = s.universe(4)
    differs = 0
    for mw, ca in itertools.product([1, 3, 5], [True, False]):
        kw = dict(check_angle=ca, additional_donors=['N'], additional_acceptors=['O'], add_donors_without_hydrogen=not ca)
        wa = ref.WireAnalysis("protein", u, **kw)
        wa.set_water_wires(max_water=mw, water_in_convex_hull=mw)
        p = hbond_port.run_wires(u, "protein", max_water=mw, water_in_convex_hull=True, **kw)
        assert_same_wires(wire_lists(p.wire_lengths), wire_lists(wa.wire_lengths), "hull %d %s" % (mw, ca))
        assert list(p.wire_lengths) == list(wa.wire_lengths)
        plain = hbond_port.run_wires(u, "protein", max_water=mw, **kw)
        differs += wire_lists(plain.wire_lengths) != wire_lists(p.wire_lengths)
    assert differs > 0            # the option changes the result on this system (else the test would prove nothing)


@pytest.mark.reference
def test_port_wires_atomwise_equal_reference():
    """residuewise=False: the port follows the reference (wire keys by the residue's first entry, direct bonds by atom)."""
    import warnings
    warnings.filterwarnings("ignore")
    ref = ref_harness.load_reference()
    s = synth.make_system(3000, 40, 5, hydrogens=True, water="TIP3", n_chains=2)
    u = s.universe(4)
    for mw, ca, hull in itertools.product([1, 3], [True, False], [False, True]):
        kw = dict(check_angle=ca, additional_donors=['N'], additional_acceptors=['O'], add_donors_without_hydrogen=not ca,
                  residuewise=False)
        wa = ref.WireAnalysis("protein", u, **kw)
        wa.set_water_wires(max_water=mw, water_in_convex_hull=hull)
        p = hbond_port.run_wires(u, "protein", max_water=mw, water_in_convex_hull=hull, **kw)
        assert_same_wires(wire_lists(p.wire_lengths), wire_lists(wa.wire_lengths), "atomwise %d %s %s" % (mw, ca, hull))
        assert list(p.wire_lengths) == list(wa.wire_lengths)
