"""Generate tests/golden/wires_hull.json by running the LIVE REFERENCE WireAnalysis.set_water_wires with
`water_in_convex_hull` set (/root/reference/cgraphs/mdhbond/waterwire.py:218-228, unmodified, through
oracle/ref_harness.py) on the coordinates of the wires_traj fixture (tests/golden/wires_traj.npz).

    python tests/golden/make_golden_wires_hull.py

cgraphs itself always sets the option (proteingraphanalyser.py:484 passes `water_in_convex_hull=max_water`), so these
are the wires a cgraphs user gets - including the reference's habit of naming the in-hull waters by the unfiltered
local-water list. Same JSON form as wires_traj.json."""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness                      # noqa: E402
from tests.helpers import load_golden_wires         # noqa: E402
from tests.golden.make_golden_wires import lengths_to_lists   # noqa: E402


def main():
    warnings.filterwarnings("ignore")
    ref = ref_harness.load_reference()
    u, _ = load_golden_wires("wires_traj")
    bbk = dict(additional_donors=['N'], additional_acceptors=['O'])
    cases = []
    for mw in (1, 3, 5):
        for ca in (True, False):
            cases.append(("protein", mw, True, dict(check_angle=ca, add_donors_without_hydrogen=not ca, **bbk)))
    cases.append(("protein", 4, False, dict(cut_angle=45., distance=3.2, **bbk)))
    cases.append(("segid PA", 5, True, dict(start=1, stop=6, step=2, **bbk)))
    cases.append(("protein", 2, True, dict()))
    runs = []
    for sel, mw, direct, kw in cases:
        wa = ref.WireAnalysis(sel, u, **kw)
        wa.set_water_wires(max_water=mw, allow_direct_bonds=direct, water_in_convex_hull=mw)   # like cgraphs' own call
        runs.append(dict(selection=sel, max_water=mw, allow_direct_bonds=direct, kwargs=kw, nb_frames=int(wa.nb_frames),
                         wire_lengths=lengths_to_lists(wa.wire_lengths),
                         graph_edges=sorted(sorted(e) for e in wa.filtered_graph.edges())))
        print("%-10s max_water=%d direct=%-5s %-70s -> %d wires" % (sel, mw, direct, kw, len(wa.wire_lengths)))
    with open(os.path.join(HERE, "wires_hull.json"), "w") as f:
        json.dump(dict(fixture="wires_hull", coordinates="wires_traj.npz", generator="tests/golden/make_golden_wires_hull.py",
                       reference="evabertalan/cgraphs mdhbond.WireAnalysis (live), water_in_convex_hull=max_water",
                       numpy=np.__version__, runs=runs), f, indent=0, separators=(",", ":"))


if __name__ == "__main__":
    main()
