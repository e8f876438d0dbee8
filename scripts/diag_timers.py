#!/usr/bin/env python
"""Diagnostic: time the stages of one device-resident pass (hb_begin / hb_submit_frames_device /
hb_finalize) for the fused and the general kernels and print every library timer."""
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
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--selection", default="protein")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--options", default="", help="comma separated name=value pairs for hb_set_option")
    ap.add_argument("--modes", default="1,0")
    ap.add_argument("--pbc", type=int, default=0, help="0 none, 1 ortho, 2 triclinic")
    args = ap.parse_args()
    import torch
    from cgraphs_b200 import build as hb_build, synth
    hb_build.build()
    from cgraphs_b200.mdhbond import HbondAnalysis, _native
    from cgraphs_b200.mdhbond.calib import cos_threshold
    from cgraphs_b200.mdhbond.topology import pack_systems
    system, cfg = synth.config(args.workload, scale=args.scale)
    nf, na = args.frames, system.n_atoms
    dev = torch.device("cuda", 0)
    d_xyz = torch.empty((nf, na, 3), dtype=torch.float32, device=dev)
    system.frames_torch(0, nf, dev, out=d_xyz)
    torch.cuda.synchronize()
    hba = HbondAnalysis(args.selection, system.universe(1), check_angle=cfg["check_angle"],
                        add_donors_without_hydrogen=cfg["add_donors_without_hydrogen"])
    top = hba._topology
    tables, _ = pack_systems([top], [top.key_ids("water_around")], [na])
    ctx = hba._context()
    ctx.set_topology(tables)
    ctx.set_params(3.5, 3.5, cos_threshold(60.0) if cfg["check_angle"] else -1.0, cfg["check_angle"], True,
                   _native.METHOD_WATER_AROUND, pbc_mode=args.pbc)
    d_box = None
    if args.pbc:
        from cgraphs_b200.mdhbond.hbond import box_rows
        d_box = torch.from_numpy(np.broadcast_to(box_rows(system.dimensions).reshape(1, 9), (nf, 9)).copy()).to(dev)
    ctx.set_option("profile", 1)
    for kv in filter(None, args.options.split(",")):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    for fused in [int(x) for x in args.modes.split(",")]:
        ctx.set_option("fused", fused)
        for rep in range(args.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.begin(nf)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            ctx.submit_device(d_xyz.data_ptr(), na, nf, d_box.data_ptr() if d_box is not None else None)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            nb, wpr = ctx.finalize()
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            tm = ctx.timers()
            out = dict(fused=fused, rep=rep, frames=nf, n_bonds=nb, wall_begin_ms=1e3 * (t1 - t0),
                       wall_submit_ms=1e3 * (t2 - t1), wall_finalize_ms=1e3 * (t3 - t2), ms_run=tm["ms_run"],
                       ms_total=tm["ms_total"], fused_frames=tm["fused_frames"], fallbacks=tm["fused_fallbacks"],
                       launches=tm["n_launches"], frames_per_s=nf / (1e-3 * tm["ms_total"]) if tm["ms_total"] else None,
                       kernels={k: round(v, 4) for k, v in tm["ms_kernel"].items() if v})
            print(json.dumps(out))


if __name__ == "__main__":
    main()
