"""Generate tests/golden/hydration_traj.json by running the LIVE REFERENCE HydrationAnalysis.set_presence_around
(/root/reference/cgraphs/mdhbond/hydration.py, unmodified, through oracle/ref_harness.py) on the system of
tests/golden/wires_traj.npz (written by make_golden_wires.py).

    python tests/golden/make_golden_hydration.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness                      # noqa: E402
from tests.helpers import load_golden_wires, results_to_lists    # noqa: E402


def main():
    ref = ref_harness.load_reference()
    u, _ = load_golden_wires()
    runs = []
    for sel, r, pr, kw in (("protein", 3.5, False, {}), ("protein", 3.5, True, {}), ("protein", 8.0, False, {}),
                           ("segid PA and name CA", 5.0, True, {}), ("protein or resid 30", 4.0, False, dict(start=1, stop=6, step=2)),
                           ("protein or resname TIP3", 3.0, False, {})):
        ha = ref.HydrationAnalysis(sel, u, **kw)
        ha.set_presence_around(r, per_residue=pr)
        runs.append(dict(selection=sel, around=r, per_residue=pr, kwargs=kw, nb_frames=int(ha.nb_frames),
                         results=results_to_lists(ha.initial_results)))
        print("%-26s around=%.1f per_residue=%-5s %-32s -> %d keys" % (sel, r, pr, kw, len(ha.initial_results)))
    with open(os.path.join(HERE, "hydration_traj.json"), "w") as f:
        json.dump(dict(fixture="hydration_traj", system="wires_traj.npz", generator="tests/golden/make_golden_hydration.py",
                       reference="evabertalan/cgraphs mdhbond.HydrationAnalysis (live)", numpy=np.__version__, runs=runs),
                  f, indent=0, separators=(",", ":"))


if __name__ == "__main__":
    main()
