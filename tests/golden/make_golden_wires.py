"""Generate tests/golden/wires_traj.{npz,json} by running the LIVE REFERENCE WireAnalysis.set_water_wires
(/root/reference/cgraphs/mdhbond/waterwire.py, unmodified, through oracle/ref_harness.py).

    python tests/golden/make_golden_wires.py

The json holds, per run, the constructor kwargs, max_water / allow_direct_bonds and the reference's `wire_lengths`
as key -> list over frames (-1 = inf = no wire, 0 = direct H-bond, k = waters in the shortest wire), keys in the
reference's dict order."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cgraphs_b200 import synth                      # noqa: E402
from cgraphs_b200.universe import Universe          # noqa: E402
from oracle import ref_harness                      # noqa: E402
from tests.golden.make_golden import save_system    # noqa: E402


def lengths_to_lists(wl):
    return {k: [(-1 if not np.isfinite(x) else int(x)) for x in v] for k, v in wl.items()}


def main():
    ref = ref_harness.load_reference()
    s = synth.make_system(1800, 24, 21, hydrogens=True, water="TIP3", n_chains=2)
    xyz = s.frames(0, 6)
    save_system("wires_traj", s, xyz)
    u = Universe(s.topology, xyz)
    bbk = dict(additional_donors=['N'], additional_acceptors=['O'])
    runs = []
    cases = []
    for mw in (1, 3, 5):
        for ca in (True, False):
            for direct in (True, False):
                cases.append(("protein", mw, direct, dict(check_angle=ca, add_donors_without_hydrogen=not ca, **bbk)))
    cases.append(("protein or resname TIP3", 2, True, dict(**bbk)))
    cases.append(("protein", 4, True, dict(cut_angle=45., distance=3.2, **bbk)))
    cases.append(("segid PA", 5, True, dict(start=1, stop=6, step=2, **bbk)))
    for sel, mw, direct, kw in cases:
        wa = ref.WireAnalysis(sel, u, **kw)
        wa.set_water_wires(max_water=mw, allow_direct_bonds=direct)
        runs.append(dict(selection=sel, max_water=mw, allow_direct_bonds=direct, kwargs=kw, nb_frames=int(wa.nb_frames),
                         wire_lengths=lengths_to_lists(wa.wire_lengths),
                         graph_edges=sorted(sorted(e) for e in wa.filtered_graph.edges())))
        print("%-26s max_water=%d direct=%-5s %-70s -> %d wires" % (sel, mw, direct, kw, len(wa.wire_lengths)))
    with open(os.path.join(HERE, "wires_traj.json"), "w") as f:
        json.dump(dict(fixture="wires_traj", generator="tests/golden/make_golden_wires.py",
                       reference="evabertalan/cgraphs mdhbond.WireAnalysis (live)", numpy=np.__version__, runs=runs),
                  f, indent=0, separators=(",", ":"))


if __name__ == "__main__":
    main()
