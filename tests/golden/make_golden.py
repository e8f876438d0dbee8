"""Generate the golden vectors under tests/golden/ by running the LIVE REFERENCE
(/root/reference/cgraphs/mdhbond, unmodified, through oracle/ref_harness.py) on small synthetic systems.

    python tests/golden/make_golden.py

Each fixture is <name>.npz (topology columns + float32 coordinates actually fed to the reference) and
<name>.json (the runs: constructor kwargs, method, and the reference's result dict as
key -> list of frame indices).  The fixtures travel to the GPU box; /root/reference does not.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cgraphs_b200 import synth                      # noqa: E402
from cgraphs_b200.universe import Universe          # noqa: E402
from cgraphs_b200.mdhbond.calib import cos_threshold  # noqa: E402
from oracle import ref_harness                      # noqa: E402
from tests.helpers import results_to_lists          # noqa: E402


def save_system(name, system, xyz):
    t = system.topology
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        names=np.asarray(t.names, dtype=str), resnames=np.asarray(t.resnames, dtype=str),
                        resids=t.resids, segids=np.asarray(t.segids, dtype=str), xyz=xyz.astype(np.float32))


def run_cases(name, system, xyz, cases):
    save_system(name, system, xyz)
    u = Universe(system.topology, xyz)
    runs = []
    for c in cases:
        kw = dict(c.get("kwargs", {}))
        hba = ref_harness.run_reference(u, c["selection"], c["method"], around=c.get("around", 3.5),
                                        exclude_backbone_backbone=c.get("excl_bb", True), **kw)
        run = dict(selection=c["selection"], method=c["method"], around=c.get("around", 3.5),
                   excl_bb=c.get("excl_bb", True), kwargs=kw, nb_frames=int(hba.nb_frames),
                   cos_threshold=float(cos_threshold(float(kw.get("cut_angle", 60.)))),
                   n_bonds=len(hba.initial_results), results=results_to_lists(hba.initial_results),
                   graph_edges=sorted(sorted(e) for e in hba.filtered_graph.edges()))
        runs.append(run)
        print("%-14s %-28s %-13s %-60s -> %d bonds" % (name, c["selection"], c["method"], kw, run["n_bonds"]))
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(dict(fixture=name, generator="tests/golden/make_golden.py", reference="evabertalan/cgraphs mdhbond (live)",
                       numpy=np.__version__, runs=runs), f, indent=0, separators=(",", ":"))


def main():
    # 1. solvated two-chain peptide trajectory (both chains share resids -> tie-resid orientation)
    s = synth.make_system(1800, 24, 21, hydrogens=True, water="TIP3", n_chains=2)
    xyz = s.frames(0, 6)
    bbk = dict(additional_donors=['N'], additional_acceptors=['O'])
    cases = []
    for sel in ("protein", "protein or resname TIP3"):
        for method in ("in_selection", "water_around"):
            for ca in (True, False):
                for rw in (True, False):
                    cases.append(dict(selection=sel, method=method, kwargs=dict(check_angle=ca, residuewise=rw, **bbk)))
    cases.append(dict(selection="protein", method="water_around", around=5.0, kwargs=dict(cut_angle=45., **bbk)))
    cases.append(dict(selection="protein", method="water_around", around=2.0, kwargs=dict(distance=3.2)))
    cases.append(dict(selection="protein", method="in_selection", excl_bb=False, kwargs=dict(**bbk)))
    cases.append(dict(selection="protein", method="water_around", kwargs=dict(start=1, stop=5, step=2)))
    cases.append(dict(selection="segid PA", method="water_around", kwargs=dict(exclude_donors=['OG'], exclude_acceptors=['OD1'])))
    run_cases("peptide_traj", s, xyz, cases)

    # 2. static crystal-like structure without hydrogens (how cgraphs calls the path)
    s = synth.make_system(1500, 60, 22, hydrogens=False, water="HOH")
    xyz = s.frames(0, 1, static=True)
    k = dict(check_angle=False, add_donors_without_hydrogen=True)
    cases = [dict(selection="protein", method=m, kwargs=dict(residuewise=rw, **k, **extra))
             for m in ("in_selection", "water_around") for rw in (True, False)
             for extra in (dict(), dict(additional_donors=['N'], additional_acceptors=['O']))]
    cases.append(dict(selection="protein or resname HOH", method="water_around", kwargs=dict(**k)))
    run_cases("crystal_static", s, xyz, cases)

    # 3. planted pairs within a few float32 ulps of the distance cut-off
    s = synth.kat_distance_system(1000, 3)
    k = dict(check_angle=False, add_donors_without_hydrogen=True)
    run_cases("kat_distance", s, s.frames(0, 1, static=True),
              [dict(selection="resname TIP3", method="in_selection", kwargs=dict(**k)),
               dict(selection="resname TIP3", method="water_around", kwargs=dict(residuewise=False, **k))])

    # 4. planted D-H...A triples within a few float32 ulps of the angle threshold
    s = synth.kat_angle_system(1001, 4, cos_threshold=cos_threshold(60.))
    run_cases("kat_angle", s, s.frames(0, 1, static=True),
              [dict(selection="resname TIP3", method="in_selection", kwargs=dict()),
               dict(selection="resname TIP3", method="water_around", kwargs=dict())])


if __name__ == "__main__":
    main()
