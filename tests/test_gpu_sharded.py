"""`shard=(rank, world)` end to end on the GPU: two processes (gloo rendezvous, both on cuda:0 - the round-end box has
one GPU) each run their frame block through libhbgpu, exchange keys / rows and must both hold exactly the result of
a single process: H-bond occupancies and water wires."""
import os
import pickle
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from cgraphs_b200 import synth
from tests.helpers import assert_same_results, assert_same_wires, wire_lists

pytestmark = pytest.mark.gpu

N_FRAMES = 70
KW = dict(additional_donors=['N'], additional_acceptors=['O'])


def _system():
    return synth.make_system(2500, 30, 9, hydrogens=True, water="TIP3", n_chains=2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cgraphs_b200.mdhbond import HbondAnalysis, WireAnalysis
        u = _system().universe(N_FRAMES)
        hba = HbondAnalysis("protein", u, device=0, shard=(rank, world), **KW)
        hba.set_hbonds_in_selection_and_water_around(3.5)
        wa = WireAnalysis("protein", u, device=0, shard=(rank, world), **KW)
        wa.set_water_wires(max_water=3)
        with open(os.path.join(out_dir, "rank%d.pkl" % rank), "wb") as f:
            pickle.dump((hba.initial_results, wire_lists(wa.wire_lengths)), f)
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_single_process(tmp_path):
    from oracle import hbond_port
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    u = _system().universe(N_FRAMES)
    ref_h = hbond_port.run(u, "protein", "water_around", around=3.5, **KW).initial_results
    ref_w = wire_lists(hbond_port.run_wires(u, "protein", max_water=3, **KW).wire_lengths)
    assert len(ref_h) > 20 and len(ref_w) > 5
    for r in range(world):
        with open(os.path.join(str(tmp_path), "rank%d.pkl" % r), "rb") as f:
            got_h, got_w = pickle.load(f)
        assert_same_results(got_h, ref_h, "rank %d" % r)
        assert_same_wires(got_w, ref_w, "rank %d" % r)
