"""Frame sharding and the bond-key merge on CPU: world_size 2 and 3 over gloo.

Each rank computes its own frame block (here with the CPU oracle standing in for the GPU search, the
merge code is the product's), the ranks exchange their distinct bond keys with one all-gather and must
all end up with exactly the matrix a single process computes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cgraphs_b200 import synth
from cgraphs_b200.mdhbond.sharding import merge_shards, shard_range

N_FRAMES = 75          # not a multiple of 32: the last shard is short, one shard may be empty


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _encode(results, ids, lo, hi):
    """result dict -> (sorted uint64 keys, packed uint32 rows over frames [lo, hi)) like hb_get_keys / hb_get_bits."""
    gid = {s: i for i, s in enumerate(ids)}
    rows = []
    for k, v in results.items():
        a, b = k.split(":")
        if v[lo:hi].any():
            rows.append(((gid[a] << 32) | gid[b], v[lo:hi]))
    rows.sort(key=lambda r: r[0])
    n_words = (max(hi - lo, 1) + 31) // 32
    keys = np.array([r[0] for r in rows], dtype=np.uint64)
    words = np.zeros((len(rows), n_words), dtype=np.uint32)
    for i, (_, bits) in enumerate(rows):
        padded = np.zeros(n_words * 32, dtype=np.uint8)
        padded[:hi - lo] = bits
        words[i] = np.packbits(padded, bitorder="little").view(np.uint32)
    return keys, words


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import hbond_port
        s = synth.make_system(2500, 25, 21)
        u = s.universe(N_FRAMES)
        pt = hbond_port.PortTopology(u, "protein")
        ids = sorted(set(pt.all_ids))
        lo, hi = shard_range(N_FRAMES, rank, world)
        local = {}
        if hi > lo:
            local = hbond_port.run_block(pt, s.frames(lo, hi), "water_around", 3.5, nb_frames=N_FRAMES, col0=lo)
        keys, words = _encode(local, ids, lo, hi)
        gkeys, gwords = merge_shards(keys, words, lo, hi, N_FRAMES, (rank, world))
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), keys=gkeys, words=gwords)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_single_process(world, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import hbond_port
    s = synth.make_system(2500, 25, 21)
    u = s.universe(N_FRAMES)
    pt = hbond_port.PortTopology(u, "protein")
    ids = sorted(set(pt.all_ids))
    full = hbond_port.run_block(pt, s.frames(0, N_FRAMES), "water_around", 3.5, nb_frames=N_FRAMES)
    keys, words = _encode(full, ids, 0, N_FRAMES)
    assert len(keys) > 20
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(got["keys"], keys), "rank %d: global key list differs" % r
        assert np.array_equal(got["words"], words), "rank %d: occupancy matrix differs" % r


def test_shard_ranges_partition_on_word_boundaries():
    for n, world in [(75, 2), (75, 3), (10000, 8), (31, 4), (1, 8), (64, 2)]:
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
        assert all(a % 32 == 0 for a, _ in spans)
