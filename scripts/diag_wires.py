#!/usr/bin/env python
"""Diagnostic / measurement for the water-wire path: device-resident frames -> hb_begin / hb_submit_frames_device /
hb_finalize in wire mode (hb_set_wires), with the per-kernel-class timers; optionally the CPU port
(oracle.hbond_port.run_wires - measurement harness only) on a few frames for the CPU baseline."""
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
    ap.add_argument("--frames", type=int, default=2048)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--selection", default="protein")
    ap.add_argument("--max-water", type=int, default=5)
    ap.add_argument("--check-angle", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=0)
    ap.add_argument("--options", default="")
    ap.add_argument("--profile", type=int, default=1)
    args = ap.parse_args()
    import torch
    from cgraphs_b200 import build as hb_build, synth
    hb_build.build()
    from cgraphs_b200.mdhbond import WireAnalysis, _native
    from cgraphs_b200.mdhbond.calib import cos_threshold
    from cgraphs_b200.mdhbond.topology import pack_systems
    system, cfg = synth.config(args.workload, scale=args.scale)
    nf, na = args.frames, system.n_atoms
    ca = bool(args.check_angle)
    kw = dict(check_angle=ca, add_donors_without_hydrogen=not ca, additional_donors=['N'], additional_acceptors=['O'])
    dev = torch.device("cuda", 0)
    d_xyz = torch.empty((nf, na, 3), dtype=torch.float32, device=dev)
    system.frames_torch(0, nf, dev, out=d_xyz)
    torch.cuda.synchronize()
    wa = WireAnalysis(args.selection, system.universe(1), **kw)
    top = wa._topology
    tables, _ = pack_systems([top], [top.key_ids("water_around")], [na])
    ctx = wa._context()
    ctx.set_topology(tables)
    ctx.set_wires(wa._wire_ent_node, wa._n_wire_nodes, args.max_water, True)
    ctx.set_option("profile", args.profile)
    for kv in filter(None, args.options.split(",")):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    around = (args.max_water + 1) * 3.5 / 2.
    ctx.set_params(3.5, around, cos_threshold(60.0) if ca else -1.0, ca, False, _native.METHOD_WATER_AROUND)
    for rep in range(args.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.begin(nf)
        ctx.submit_device(d_xyz.data_ptr(), na, nf, None)
        nb, wpr = ctx.finalize()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        tm = ctx.timers()
        print(json.dumps(dict(rep=rep, frames=nf, n_res=wa._n_wire_nodes, n_pairs=nb, wpr=wpr, wall_ms=1e3 * (t1 - t0),
                              ms_run=tm["ms_run"], ms_total=tm["ms_total"], fused_frames=tm["fused_frames"],
                              fallbacks=tm["fused_fallbacks"], launches=tm["n_launches"], edges_per_frame=tm["pairs_emitted"] / nf,
                              frames_per_s=nf / (1e-3 * tm["ms_run"]) if tm["ms_run"] else None,
                              kernels={k: round(v, 3) for k, v in tm["ms_kernel"].items() if v})))
    if args.cpu_frames:
        from oracle import hbond_port
        u = system.universe(args.cpu_frames)
        t0 = time.perf_counter()
        r = hbond_port.run_wires(u, args.selection, max_water=args.max_water, **kw)
        dt = time.perf_counter() - t0
        print(json.dumps(dict(cpu_port_frames=args.cpu_frames, seconds=dt, frames_per_s=args.cpu_frames / dt,
                              n_pairs=len(r.wire_lengths))))


if __name__ == "__main__":
    main()
