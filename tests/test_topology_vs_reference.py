"""The product's vectorised topology compiler reproduces the tables of the reference constructor
(cgraphs/mdhbond/network.py:40-97) - checked against the live reference object."""
import itertools

import numpy as np
import pytest

from cgraphs_b200 import synth
from cgraphs_b200.mdhbond.topology import AtomColumns, CompiledTopology, WATER_DEFINITION
from oracle import ref_harness, hbond_port

pytestmark = pytest.mark.reference

CASES = [
    dict(hydrogens=True, water="TIP3", kw=dict()),
    dict(hydrogens=True, water="TIP3", kw=dict(residuewise=False, additional_donors=['N'], additional_acceptors=['O'])),
    dict(hydrogens=False, water="HOH", kw=dict(check_angle=False, add_donors_without_hydrogen=True)),
    dict(hydrogens=True, water="TIP3", kw=dict(exclude_donors=['OG'], exclude_acceptors=['OD1'])),
]


@pytest.mark.parametrize("case,selection", list(itertools.product(range(len(CASES)), ["protein", "protein or resname TIP3 HOH"])))
def test_tables_match_reference(case, selection):
    c = CASES[case]
    s = synth.make_system(2500, 30, 11 + case, hydrogens=c["hydrogens"], water=c["water"], n_chains=2)
    u = s.universe(2)
    ref = ref_harness.load_reference()
    hba = ref.HbondAnalysis(selection, u, **c["kw"])
    cols = AtomColumns(u)
    kw = {k: v for k, v in c["kw"].items()}
    top = CompiledTopology(cols, u.select_atoms(selection).indices, u.select_atoms(WATER_DEFINITION).indices,
                           u.trajectory.frames.frame(0), **kw)
    assert top.nD == hba._nb_donors and top.nA == hba._nb_acceptors
    if top.nD:
        assert np.array_equal(top.donors, hba._donors.indices)
    if top.nA:
        assert np.array_equal(top.acceptors, hba._acceptors.indices)
    assert np.array_equal(top.water_atoms, hba._water.indices)
    assert top.first_water == hba._first_water_id
    assert top.first_water_h == hba._first_water_hydrogen_id
    if len(top.h_atom):
        assert np.array_equal(top.h_atom, hba._hydrogen.indices)
    mine = [list(map(int, l)) for l in top.heavy2hydrogen()]
    theirs = [list(map(int, l)) for l in hba.heavy2hydrogen]
    # without water hydrogens the reference list simply stops after the selection entries (network.py:86)
    assert mine[:len(theirs)] == theirs and all(l == [] for l in mine[len(theirs):])
    assert top.all_ids == hba._all_ids
    assert top.all_ids_atomwise == hba._all_ids_atomwise
    assert np.array_equal(top.resids, hba._resids)
    # and the oracle port derives the same
    pt = hbond_port.PortTopology(u, selection, **kw)
    assert np.array_equal(pt.entry_atom, top.entry_atom)
    assert pt.heavy2hydrogen == mine[:len(pt.heavy2hydrogen)]
    assert pt.all_ids == top.all_ids and pt.all_ids_atomwise == top.all_ids_atomwise


def test_frame_zero_used_for_hydrogen_detection():
    """A.1 step 5: donor-hydrogen distances are measured on frame 0 even when start > 0."""
    s = synth.make_system(1200, 12, 5)
    u = s.universe(6)
    ref = ref_harness.load_reference()
    hba = ref.HbondAnalysis("protein", u, start=3)
    cols = AtomColumns(u)
    top = CompiledTopology(cols, u.select_atoms("protein").indices, u.select_atoms(WATER_DEFINITION).indices,
                           u.trajectory.frames.frame(0))
    assert np.array_equal(top.donors, hba._donors.indices)
    assert hba.nb_frames == 3
