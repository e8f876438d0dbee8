#!/usr/bin/env python
"""Diagnostic: end-to-end frames/s through hb_submit_frames from pinned and from pageable host memory
(the latter is what the drop-in class submits: numpy blocks), with 1 and 8 staging threads."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4096)
    args = ap.parse_args()
    import torch
    from cgraphs_b200 import build as hb_build, synth
    hb_build.build()
    from cgraphs_b200.mdhbond import HbondAnalysis, _native
    from cgraphs_b200.mdhbond.calib import cos_threshold
    from cgraphs_b200.mdhbond.topology import pack_systems
    system, cfg = synth.config("C3")
    nf, na = args.frames, system.n_atoms
    dev = torch.device("cuda", 0)
    d = torch.empty((nf, na, 3), dtype=torch.float32, device=dev)
    system.frames_torch(0, nf, dev, out=d)
    pageable = d.cpu().numpy()
    pin = _native.PinnedBuffer((nf, na, 3))
    pin.array[...] = pageable
    hba = HbondAnalysis("protein", system.universe(1))
    top = hba._topology
    tables, _ = pack_systems([top], [top.key_ids("water_around")], [na])
    ctx = hba._context()
    ctx.set_topology(tables)
    ctx.set_params(3.5, 3.5, cos_threshold(60.0), True, True, _native.METHOD_WATER_AROUND)
    for name, ptr, threads in (("pinned", pin.ptr, 8), ("pageable, 1 staging thread", pageable.ctypes.data, 1),
                               ("pageable, 8 staging threads", pageable.ctypes.data, 8),
                               ("pageable, 16 staging threads", pageable.ctypes.data, 16)):
        ctx.set_option("stage_threads", threads)
        best = None
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.begin(nf)
            ctx.submit_raw(ptr, na, nf)
            keys, words = ctx.results()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(json.dumps(dict(source=name, frames_per_s=nf / best, gb_per_s=ctx.timers()["h2d_bytes"] / best / 1e9, n_bonds=len(keys))))


if __name__ == "__main__":
    main()
