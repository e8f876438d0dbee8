"""Pin the oracle: the CPU restatement equals the reference's own HbondAnalysis (run live through
oracle/ref_harness.py) on random systems, for both methods and all criteria combinations."""
import itertools

import pytest

from cgraphs_b200 import synth
from cgraphs_b200.mdhbond.calib import cos_threshold
from oracle import hbond_port, ref_harness
from tests.helpers import assert_same_results, graph_edges

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("seed", [0, 1])
def test_port_equals_reference_random(seed):
    s = synth.make_system(3000, 36, 100 + seed, hydrogens=True, water="TIP3", n_chains=2)
    u = s.universe(3)
    for sel, method, ca, rw in itertools.product(["protein", "protein or resname TIP3"],
                                                 ["in_selection", "water_around"], [True, False], [True, False]):
        kw = dict(check_angle=ca, residuewise=rw, additional_donors=['N'], additional_acceptors=['O'])
        ref = ref_harness.run_reference(u, sel, method, **kw)
        for backend, cthr in (("kdtree", None), ("brute", cos_threshold(60.))):
            port = hbond_port.run(u, sel, method, backend=backend, cos_threshold=cthr, **kw)
            assert_same_results(port.initial_results, ref.initial_results, "%s %s %s %s %s" % (sel, method, ca, rw, backend))
            assert graph_edges(port.filtered_graph) == graph_edges(ref.filtered_graph)
            assert port.initial_results is port.filtered_results


def test_port_equals_reference_membrane_like():
    s, cfg = synth.config("C3", scale=0.1)
    u = s.universe(2)
    ref = ref_harness.run_reference(u, "protein", "water_around", around=3.5)
    port = hbond_port.run(u, "protein", "water_around", around=3.5)
    assert len(ref.initial_results) > 3
    assert_same_results(port.initial_results, ref.initial_results)


def test_port_equals_reference_distance_only_static():
    s, cfg = synth.config("C1", scale=0.5)
    u = s.universe(1, static=True)
    kw = dict(check_angle=False, add_donors_without_hydrogen=True)
    for method in ("in_selection", "water_around"):
        ref = ref_harness.run_reference(u, "protein", method, **kw)
        port = hbond_port.run(u, "protein", method, **kw)
        assert_same_results(port.initial_results, ref.initial_results, method)


def test_second_call_replaces_first_result():
    """basic.py:94-96: cgraphs calls in_selection then water_around; the second result replaces the first."""
    s, cfg = synth.config("C1", scale=0.3)
    u = s.universe(1, static=True)
    ref = ref_harness.load_reference()
    hba = ref.HbondAnalysis("protein", u, check_angle=False, add_donors_without_hydrogen=True)
    hba.set_hbonds_in_selection()
    first = dict(hba.initial_results)
    hba.set_hbonds_in_selection_and_water_around(3.5)
    port = hbond_port.run(u, "protein", "water_around", check_angle=False, add_donors_without_hydrogen=True)
    assert_same_results(port.initial_results, hba.initial_results)
    assert set(first) <= set(hba.initial_results)


@pytest.mark.parametrize("residuewise", [True, False])
def test_port_ion_contacts_equal_reference(residuewise):
    """set_add_ion_interactions (hbond.py:57-94), called at the end of both set_hbonds_* methods when ions are given."""
    from tests.helpers import ion_system, port_with_ions
    u = ion_system()
    for method in ("in_selection", "water_around"):
        ref = ref_harness.run_reference(u, "protein", method, ions=['SOD', 'CLA'], residuewise=residuewise)
        got = port_with_ions(u, "protein", method, ['SOD', 'CLA'], residuewise=residuewise)
        assert {k: v.tobytes() for k, v in ref.initial_results.items()} == {k: v.tobytes() for k, v in got.items()}
        assert any(k.split(':')[1].split('-')[1] in ('SOD', 'CLA') for k in got)
