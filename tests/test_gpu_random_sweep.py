"""Randomised parity sweep: small random systems x random parameters, CUDA path (fused and general kernels,
with and without minimum image) against the CPU oracle. Deterministic (seeded numpy generator), 100 cases."""
import numpy as np
import pytest

from cgraphs_b200 import synth
from tests.helpers import assert_same_results

pytestmark = pytest.mark.gpu

SELECTIONS = ["protein", "protein or resname TIP3", "resname TIP3", "protein and not name N O"]


def _case(rng):
    return dict(
        n_atoms=int(rng.integers(700, 3200)), n_res=int(rng.integers(4, 26)), seed=int(rng.integers(0, 10 ** 6)),
        kind=["ortho", "tric"][int(rng.integers(2))], n_frames=int(rng.integers(1, 6)),
        selection=SELECTIONS[int(rng.integers(len(SELECTIONS)))],
        method=["in_selection", "water_around"][int(rng.integers(2))],
        check_angle=bool(rng.integers(2)), residuewise=bool(rng.integers(2)),
        distance=float(np.round(rng.uniform(2.6, 4.4), 2)), around=float(np.round(rng.uniform(1.0, 7.5), 2)),
        cut_angle=float(np.round(rng.uniform(15.0, 110.0), 1)), excl_bb=bool(rng.integers(2)),
        pbc=bool(rng.integers(2)), fused=int(rng.integers(2)), extra=bool(rng.integers(2)), frame_pairs=int(rng.integers(2)))


@pytest.mark.parametrize("block", range(10))
def test_random_parameter_sweep(block):
    from cgraphs_b200.mdhbond import HbondAnalysis
    from oracle import hbond_port
    rng = np.random.default_rng(9100 + block)
    done = 0
    while done < 10:
        c = _case(rng)
        s = synth.make_system(c["n_atoms"], c["n_res"], c["seed"], dimensions_kind=c["kind"])
        u = s.universe(c["n_frames"])
        kw = dict(check_angle=c["check_angle"], residuewise=c["residuewise"], distance=c["distance"], cut_angle=c["cut_angle"])
        if c["extra"]:
            kw.update(additional_donors=['N'], additional_acceptors=['O'])
        try:
            ref = hbond_port.run(u, c["selection"], c["method"], around=c["around"], exclude_backbone_backbone=c["excl_bb"],
                                 **kw).initial_results
        except AssertionError:          # e.g. a selection without donors: the product must raise the same way
            with pytest.raises(AssertionError):
                h = HbondAnalysis(c["selection"], u, **kw)
                getattr(h, "set_hbonds_in_selection")(c["excl_bb"]) if c["method"] == "in_selection" else \
                    h.set_hbonds_in_selection_and_water_around(c["around"], c["excl_bb"])
            continue
        edge = min(hbond_port.box_rows(s.dimensions).diagonal())
        if 2.0 * max(c["distance"], c["around"]) >= 0.98 * edge:
            c["pbc"] = False                 # cut-off too close to half the cell for the minimum-image mode
        h = HbondAnalysis(c["selection"], u, pbc=c["pbc"], native_options={"fused": c["fused"], "frame_pairs": c["frame_pairs"]}, **kw)
        if c["method"] == "in_selection":
            h.set_hbonds_in_selection(c["excl_bb"])
        else:
            h.set_hbonds_in_selection_and_water_around(c["around"], c["excl_bb"])
        assert_same_results(h.initial_results, ref, str(c))     # skin systems: minimum image == reference semantics
        h.close()
        done += 1
