#!/usr/bin/env python
"""Small end-to-end runs of every kernel path for compute-sanitizer (memcheck / racecheck / synccheck):
fused and general kernels, water_around and in_selection, angle on and off, minimum-image mode, structure batch,
finalisation with both sort paths, the key-merge kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from cgraphs_b200 import synth
    from cgraphs_b200.mdhbond import HbondAnalysis, HbondBatch, _native
    s = synth.make_system(2500, 20, 4, dimensions_kind="tric")
    u = s.universe(5)
    n = 0
    for opts in ({"fused": 1}, {"fused": 0}, {"fused": 0, "frame_pairs": 0}, {"fused": 0, "frame_pairs": 0, "pair_buf_bytes": 0},
                 {"fused": 1, "force_radix_sort": 1}):
        for sel, method, ca, pbc in [("protein", "water_around", True, False), ("protein or resname TIP3", "water_around", True, False),
                                     ("protein", "in_selection", False, False), ("protein", "water_around", True, True)]:
            h = HbondAnalysis(sel, u, check_angle=ca, native_options=opts, pbc=pbc,
                              additional_donors=['N'], additional_acceptors=['O'])
            if method == "in_selection":
                h.set_hbonds_in_selection()
            else:
                h.set_hbonds_in_selection_and_water_around(3.5)
            n += len(h.initial_results)
            h.compute_joint_occupancy()
            h.filter_occupancy(0.3)
            h._compute_graph_in_frame(2, use_filtered=False)
            h.close()
    systems = synth.c5_structures(4, seed0=5200)
    b = HbondBatch([x.universe(1, static=True) for x in systems], "protein", check_angle=False, add_donors_without_hydrogen=True)
    b.set_hbonds_in_selection_and_water_around(3.5)
    b.get_conserved_graph(0.5)
    ctx = _native.Context(0)
    g = torch.zeros((3, 9), dtype=torch.int64, device="cuda")
    g[:, 0] = torch.tensor([3, 8, 0])
    g[0, 1:4] = torch.tensor([5, 7, 9]); g[1, 1:9] = torch.arange(2, 18, 2)
    out = torch.empty(24, dtype=torch.int64, device="cuda"); info = torch.empty(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.merge_keys_device(g.data_ptr(), 3, 8, out.data_ptr(), info.data_ptr())
    ctx.sync()
    print("sanitize run ok:", n, "bonds,", info.tolist())


if __name__ == "__main__":
    main()
