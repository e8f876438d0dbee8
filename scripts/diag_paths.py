#!/usr/bin/env python
"""Diagnostic: the device-resident pass and the host-frame pass of the C3 workload must give the same keys and occupancy
rows; prints where they differ (bond, frames) and which of the two agrees with the CPU oracle there.
usage: diag_paths.py [frames] [name=value options ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from cgraphs_b200 import build as hb_build, synth
    hb_build.build()
    from cgraphs_b200.mdhbond import HbondAnalysis, _native
    from cgraphs_b200.mdhbond.calib import cos_threshold
    from cgraphs_b200.mdhbond.topology import pack_systems
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    system, cfg = synth.config("C3")
    na = system.n_atoms
    dev = torch.device("cuda", 0)
    d_xyz = torch.empty((nf, na, 3), dtype=torch.float32, device=dev)
    system.frames_torch(0, nf, dev, out=d_xyz)
    torch.cuda.synchronize()
    hba = HbondAnalysis("protein", system.universe(1), check_angle=True, add_donors_without_hydrogen=False)
    top = hba._topology
    tables, _ = pack_systems([top], [top.key_ids("water_around")], [na])
    ctx = hba._context()
    ctx.set_topology(tables)
    ctx.set_params(3.5, 3.5, cos_threshold(60.0), True, True, _native.METHOD_WATER_AROUND)
    for kv in sys.argv[2:]:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))

    def dev_pass():
        ctx.begin(nf)
        ctx.submit_device(d_xyz.data_ptr(), na, nf)
        ctx.finalize()
        return ctx.results()

    pin = _native.PinnedBuffer((nf, na, 3))
    host = pin.array
    chunk = max(1, (1 << 30) // (na * 12))
    for s in range(0, nf, chunk):
        host[s:s + chunk] = d_xyz[s:min(s + chunk, nf)].cpu().numpy()

    def host_pass():
        ctx.begin(nf)
        ctx.submit_raw(host.ctypes.data, na, nf)
        ctx.finalize()
        return ctx.results()

    runs = [("device", dev_pass()), ("device again", dev_pass()), ("host", host_pass()), ("host again", host_pass())]
    k0, w0 = runs[0][1]
    for name, (k, w) in runs[1:]:
        same_k = np.array_equal(k, k0)
        print("%-13s keys %d (equal to device: %s) rows equal: %s" % (name, len(k), same_k, same_k and np.array_equal(w, w0)))
        if same_k and not np.array_equal(w, w0):
            d = np.bitwise_xor(w, w0)
            rows = np.nonzero(d.any(axis=1))[0]
            frames = sorted({int(c) * 32 + b for r in rows for c in np.nonzero(d[r])[0] for b in range(32) if (int(d[r, c]) >> b) & 1})
            print("   rows differing: %d, frames differing: %d, first frames %s" % (len(rows), len(frames), frames[:20]))
            print("   timers: batches %s" % ctx.timers().get("n_batches"))
        elif not same_k:
            print("   only in this run: %d, only in device: %d" % (len(np.setdiff1d(k, k0)), len(np.setdiff1d(k0, k))))


if __name__ == "__main__":
    main()
