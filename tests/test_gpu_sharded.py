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


# ---- the key exchange inside the C ABI (hb_comm_init / hb_exchange_keys: the library's own NCCL all-gather) -----------
def _abi_worker(rank, world, out_dir):
    """No torch.distributed: the 128-byte NCCL id travels through a file, like any non-Python host would pass it."""
    import time
    import torch
    from cgraphs_b200.mdhbond import HbondAnalysis, _native
    from cgraphs_b200.mdhbond.sharding import shard_range
    id_file = os.path.join(out_dir, "nccl_id.bin")
    if rank == 0:
        with open(id_file + ".tmp", "wb") as f:
            f.write(_native.comm_unique_id())
        os.replace(id_file + ".tmp", id_file)
    while not os.path.exists(id_file):
        time.sleep(0.05)
    uid = open(id_file, "rb").read()
    torch.cuda.set_device(rank)
    u = _system().universe(N_FRAMES)
    hba = HbondAnalysis("protein", u, device=rank, **KW)
    lo, hi = shard_range(N_FRAMES, rank, world)
    hba._frame_indices = range(lo, hi)                 # this rank's frame block
    hba.nb_frames = hi - lo
    ctx = hba._context()
    ctx.comm_init(uid, world, rank)
    hba.set_hbonds_in_selection_and_water_around(3.5)   # leaves the finalised run on the context
    cap = 2048
    out = torch.zeros(world * cap, dtype=torch.int64, device="cuda:%d" % rank)
    info = torch.zeros(2, dtype=torch.int64, device="cuda:%d" % rank)
    ctx.exchange_keys(cap, out.data_ptr(), info.data_ptr())
    ctx.sync()
    n, over = info.tolist()
    with open(os.path.join(out_dir, "abi_rank%d.pkl" % rank), "wb") as f:
        pickle.dump((out[:n].cpu().numpy().view(np.uint64), over, sorted(hba.initial_results)), f)


def test_c_abi_key_exchange_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL does not put two ranks on one device)")
    from oracle import hbond_port
    world = 2
    mp.spawn(_abi_worker, args=(world, str(tmp_path)), nprocs=world, join=True)
    u = _system().universe(N_FRAMES)
    ref = hbond_port.run(u, "protein", "water_around", around=3.5, **KW).initial_results
    res = [pickle.load(open(os.path.join(str(tmp_path), "abi_rank%d.pkl" % r), "rb")) for r in range(world)]
    assert res[0][1] == 0 and res[1][1] == 0
    assert np.array_equal(res[0][0], res[1][0])                      # every rank holds the same global key list
    assert len(res[0][0]) == len(ref)                                # = the distinct bonds of the whole trajectory
    assert set(res[0][2]) | set(res[1][2]) == set(ref)
    assert np.all(np.diff(res[0][0].astype(np.int64)) > 0)
