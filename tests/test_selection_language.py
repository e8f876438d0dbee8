"""The built-in selection mini-language of cgraphs_b200.universe (used when MDAnalysis is not installed: PDB / PSF + DCD
input, the tests): keywords follow MDAnalysis' documented meaning, checked against direct numpy expressions."""
import numpy as np
import pytest

from cgraphs_b200 import synth
from cgraphs_b200.universe import PROTEIN_RESNAMES


@pytest.fixture(scope="module")
def u():
    return synth.make_system(3000, 36, 100, hydrogens=True, water="TIP3", n_chains=2).universe(2)


def _idx(u, sel):
    return np.asarray(u.select_atoms(sel).indices)


def test_keywords(u):
    t = u.topology
    prot = np.array([r in PROTEIN_RESNAMES for r in t.resnames])
    names = np.asarray(t.names, dtype=object)
    assert np.array_equal(_idx(u, "protein"), np.nonzero(prot)[0])
    assert np.array_equal(_idx(u, "backbone"), np.nonzero(prot & np.isin(names, ["N", "CA", "C", "O"]))[0])
    assert np.array_equal(_idx(u, "water"), np.nonzero(np.asarray(t.resnames, dtype=object) == "TIP3")[0])
    assert len(_idx(u, "nucleic")) == 0
    assert np.array_equal(_idx(u, "index 0:10"), np.arange(11)) and np.array_equal(_idx(u, "bynum 1:10"), np.arange(10))
    assert np.array_equal(_idx(u, "index 5"), [5]) and np.array_equal(_idx(u, "bynum 5"), [4])
    for a, b in (("resid 1 to 5", "resid 1:5"), ("resid 1-5", "resid 1:5"), ("resid 2 to 3 9", "resid 2 3 9"),
                 ("protein and not backbone", "protein and not (name N CA C O)"), ("chainID PA", "segid PA")):
        assert np.array_equal(_idx(u, a), _idx(u, b)), a


def test_byres_and_around(u):
    t = u.topology
    oh2 = _idx(u, "name OH2")
    assert np.array_equal(_idx(u, "byres name OH2"), _idx(u, "water"))
    one = _idx(u, "byres (resid 3 and segid PA and name CA)")
    assert len(one) > 1 and len(set(t.resindices[one])) == 1
    pos = np.asarray(u.trajectory.frames.frame(0), dtype=np.float64)
    prot = _idx(u, "protein")
    d2 = ((pos[:, None, :] - pos[None, prot, :]) ** 2).sum(-1).min(axis=1)
    want = np.nonzero((d2 <= 5.0 ** 2))[0]
    want = np.setdiff1d(want, prot)                     # the reference selection itself is not part of `around`
    got = _idx(u, "around 5 protein")
    assert len(np.setxor1d(got, want)) <= 2              # (float32 tree vs float64 brute force at the radius)
    near_water = _idx(u, "(around 3.5 protein) and name OH2")
    assert len(near_water) > 10 and set(near_water) <= set(oh2)
    assert set(_idx(u, "protein or byres (around 3.5 protein)")) >= set(prot)


def test_errors(u):
    for bad in ("", "resid x", "index a", "frobnicate 3", "protein and", "(protein"):
        with pytest.raises(ValueError):
            u.select_atoms(bad)


def test_precedence_follows_mdanalysis(u):
    """MDAnalysis' SelectionParser: `and` / `or` share one precedence and associate to the left, `not` binds
    tighter, `byres` / `around` take the whole following expression (core/selection.py, precedence table)."""
    same = [
        ("name CA or name N and resid 3", "(name CA or name N) and resid 3"),
        ("resid 3 and name CA or name N", "(resid 3 and name CA) or name N"),
        ("name CA or name N and resid 3 or name OH2", "((name CA or name N) and resid 3) or name OH2"),
        ("not name CA and resid 3", "(not name CA) and resid 3"),
        ("not name CA or resid 3", "(not name CA) or resid 3"),
        ("byres name CA and resid 3", "byres (name CA and resid 3)"),
        ("byres name CA and resid 3 or resid 4 and name N", "byres (((name CA and resid 3) or resid 4) and name N)"),
        ("around 3.5 protein and name CA", "around 3.5 (protein and name CA)"),
        ("name OH2 and around 3.5 protein or name CA", "name OH2 and (around 3.5 (protein or name CA))"),
        ("name OH2 and not around 3.5 protein and resid 1 to 40", "name OH2 and (not (around 3.5 (protein and resid 1:40)))"),
        ("not byres name CA and resid 3", "not (byres (name CA and resid 3))"),
    ]
    for a, b in same:
        assert np.array_equal(_idx(u, a), _idx(u, b)), a
    # and the two readings really differ on this system (the test would be vacuous otherwise)
    assert not np.array_equal(_idx(u, "name CA or name N and resid 3"), _idx(u, "name CA or (name N and resid 3)"))
    assert not np.array_equal(_idx(u, "around 3.5 protein and name CA"), _idx(u, "(around 3.5 protein) and name CA"))
