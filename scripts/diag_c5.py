#!/usr/bin/env python
"""Config C5: the conserved-network batch - N static structures searched in one batched launch
(distance-only criterion). Timing only; parity of the batch path is covered by tests/ (the oracle is not
imported outside tests/, smoke() and bench.py's CPU legs)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--structures", type=int, default=500)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    from cgraphs_b200 import build as hb_build, synth
    hb_build.build()
    from cgraphs_b200.mdhbond import HbondBatch
    t0 = time.perf_counter()
    systems = synth.c5_structures(args.structures)
    unis = [s.universe(1, static=True) for s in systems]
    t_gen = time.perf_counter() - t0
    kw = dict(check_angle=False, add_donors_without_hydrogen=True)
    t0 = time.perf_counter()
    batch = HbondBatch(unis, "protein", **kw)
    t_ctor = time.perf_counter() - t0
    out = []
    for rep in range(args.reps):
        t0 = time.perf_counter()
        batch.set_hbonds_in_selection_and_water_around(3.5)
        wall = time.perf_counter() - t0
        st = batch.last_run_stats
        out.append(dict(rep=rep, wall_s=wall, ms_run=st["ms_run"], ms_kernels=st["ms_total"], launches=st["n_launches"],
                        fused_frames=st["fused_frames"], n_bonds=len(batch.bond_keys)))
    print(json.dumps(dict(structures=len(unis), atoms=int(sum(s.n_atoms for s in systems)), gen_s=t_gen, ctor_s=t_ctor,
                          runs=out, device_structures_per_s=len(unis) / (out[-1]["ms_run"] * 1e-3))))


if __name__ == "__main__":
    main()
