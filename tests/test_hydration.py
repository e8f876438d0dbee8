"""HydrationAnalysis.set_presence_around (hydration.py:71-107): the oracle against the live
reference (CPU), the GPU class against the oracle."""
import numpy as np
import pytest

from cgraphs_b200 import synth

CASES = [("protein", 3.5, {}), ("protein", 6.0, {}), ("segid PA and name CA", 5.0, {}), ("protein or resid 40", 4.0, {}),
         ("protein", 9.5, dict(start=1, stop=5, step=2))]


def _system():
    return synth.make_system(3000, 36, 100, hydrogens=True, water="TIP3", n_chains=2)


def _same(a, b, what=""):
    ka = {k: np.asarray(v, bool).tobytes() for k, v in a.items()}
    kb = {k: np.asarray(v, bool).tobytes() for k, v in b.items()}
    assert ka == kb, "%s: %d vs %d waters, %d differ" % (what, len(ka), len(kb), sum(ka.get(k) != kb.get(k) for k in set(ka) | set(kb)))
    assert list(a) == list(b), what + ": key order"


@pytest.mark.reference
def test_port_presence_equals_reference():
    from oracle import hbond_port, ref_harness
    ref = ref_harness.load_reference()
    u = _system().universe(6)
    for sel, r, kw in CASES:
        ha = ref.HydrationAnalysis(sel, u, **kw)
        ha.set_presence_around(r)
        assert len(ha.initial_results) > 10
        for backend in ("kdtree", "brute"):
            _same(hbond_port.run_presence_around(u, sel, r, backend=backend, **kw), ha.initial_results, "%s %s" % (sel, r))


@pytest.mark.gpu
def test_gpu_presence_equals_oracle():
    from cgraphs_b200.mdhbond import HydrationAnalysis
    from oracle import hbond_port
    u = _system().universe(6)
    for sel, r, kw in CASES:
        ref = hbond_port.run_presence_around(u, sel, r, **kw)
        for options in ({}, {"lw_smem": 0}, {"table_capacity": 16}):
            ha = HydrationAnalysis(sel, u, native_options=options, **kw)
            ha.set_presence_around(r)
            _same(ha.initial_results, ref, "%s %s %s" % (sel, r, options))
            assert ha.initial_results is ha.filtered_results
    ha.filter_occupancy(0.5)
    assert set(ha.filtered_results) == {k for k, v in ref.items() if v.mean() > 0.5}


@pytest.mark.gpu
def test_gpu_presence_membrane_like():
    from cgraphs_b200.mdhbond import HydrationAnalysis
    from oracle import hbond_port
    s, cfg = synth.config("C3", scale=0.3)
    u = s.universe(40)
    ref = hbond_port.run_presence_around(u, "protein", 5.0)
    ha = HydrationAnalysis("protein", u)
    ha.set_presence_around(5.0)
    assert len(ref) > 100
    _same(ha.initial_results, ref)


@pytest.mark.reference
def test_port_presence_per_residue_equals_reference():
    from oracle import hbond_port, ref_harness
    ref = ref_harness.load_reference()
    u = _system().universe(4)
    for sel, r in (("protein", 3.5), ("segid PA and name CA", 5.0), ("protein or resid 40", 4.0)):
        ha = ref.HydrationAnalysis(sel, u)
        ha.set_presence_around(r, per_residue=True)
        got = hbond_port.run_presence_around(u, sel, r, per_residue=True)
        assert {k: v.tobytes() for k, v in got.items()} == {k: v.tobytes() for k, v in ha.initial_results.items()}


@pytest.mark.gpu
def test_gpu_presence_per_residue_equals_oracle():
    """per_residue=True: 'residue id:water id' keys from the distance-only pair search (selected atoms as entries)."""
    from cgraphs_b200.mdhbond import HydrationAnalysis
    from oracle import hbond_port
    u = _system().universe(6)
    for sel, r, kw in (("protein", 3.5, {}), ("segid PA and name CA", 5.0, {}), ("protein", 7.25, dict(start=1, stop=6, step=2))):
        ref = hbond_port.run_presence_around(u, sel, r, per_residue=True, **kw)
        assert len(ref) > 20
        for options in ({}, {"fused": 0}, {"fused": 0, "frame_pairs": 0}):
            ha = HydrationAnalysis(sel, u, native_options=options, **kw)
            ha.set_presence_around(r, per_residue=True)
            assert {k: v.tobytes() for k, v in ha.initial_results.items()} == {k: v.tobytes() for k, v in ref.items()}, (sel, r, options)
    with pytest.raises(NotImplementedError):
        HydrationAnalysis("protein or resname TIP3", u).set_presence_around(3.5, per_residue=True)


def _golden_runs():
    import json
    import os
    from tests.helpers import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "hydration_traj.json")) as f:
        return json.load(f)["runs"]


@pytest.mark.parametrize("i", range(6))
def test_port_presence_equals_golden(i):
    """Golden vectors of the live reference (tests/golden/make_golden_hydration.py) - travel to the GPU box."""
    from oracle import hbond_port
    from tests.helpers import lists_to_results, load_golden_wires
    u, _ = load_golden_wires()
    r = _golden_runs()[i]
    got = hbond_port.run_presence_around(u, r["selection"], r["around"], per_residue=r["per_residue"], **r["kwargs"])
    want = lists_to_results(r["results"], r["nb_frames"])
    assert {k: v.tobytes() for k, v in got.items()} == {k: v.tobytes() for k, v in want.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(6))
def test_gpu_presence_equals_golden(i):
    from cgraphs_b200.mdhbond import HydrationAnalysis
    from tests.helpers import lists_to_results, load_golden_wires
    u, _ = load_golden_wires()
    r = _golden_runs()[i]
    ha = HydrationAnalysis(r["selection"], u, **r["kwargs"])
    ha.set_presence_around(r["around"], per_residue=r["per_residue"])
    want = lists_to_results(r["results"], r["nb_frames"])
    assert {k: v.tobytes() for k, v in ha.initial_results.items()} == {k: v.tobytes() for k, v in want.items()}
