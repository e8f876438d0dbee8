"""Host side of the water-wire path without a GPU: the row decoding of `cgraphs_b200.mdhbond.waterwire`
(key orientation by first discovery) and the shard merge of `sharding.merge_wire_shards` over gloo (world 2 and 3).
The per-frame search + BFS is stood in for by the oracle, encoded exactly as `hb_set_wires` documents its rows
(include/hbgpu.h); decoding and merging are the product's."""
import os
import pickle
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from cgraphs_b200 import synth
from cgraphs_b200.mdhbond.sharding import merge_wire_shards, shard_range
from cgraphs_b200.mdhbond.waterwire import decode_wire_rows, split_wire_rows
from tests.helpers import assert_same_wires, wire_lists

N_FRAMES = 70
MAX_WATER = 3
KW = dict(check_angle=False, add_donors_without_hydrogen=True, additional_donors=['N'], additional_acceptors=['O'])


def _system():
    return synth.make_system(2500, 30, 9, hydrogens=True, water="TIP3", n_chains=2)


def _device_rows(u, lo, hi):
    """What a wire run over frames [lo, hi) leaves in the table: (keys, uint32 rows [n, nb4 + 2], node_entry, ids)."""
    from oracle import hbond_port
    pt = hbond_port.PortTopology(u, "protein", **KW)
    first = hbond_port.wire_node_of_entry(pt)
    node_entry = np.unique(first)
    node = np.searchsorted(node_entry, first)               # dense residue node of every selection entry
    n = hi - lo
    nb4 = (((n + 3) // 4) + 1) & ~1
    rows = {}

    def row(a, b):
        return rows.setdefault((min(a, b), max(a, b)), [np.zeros(nb4 * 4, np.uint8), np.uint64(0)])

    frames = u.trajectory.frames
    for col, f in enumerate(range(lo, hi)):
        pos = np.asarray(frames.frame(f), dtype=np.float32)
        edges = hbond_port.wire_frame_edges(pt, pos, MAX_WATER)
        wires, directs = hbond_port.wires_of_frame(pt, first, edges, MAX_WATER, True)
        for (s, t), k in wires.items():
            row(node[s], node[t])[0][col] |= k
        for d_ent, a_ent in directs:
            nd, na = node[d_ent], node[a_ent]
            r = row(nd, na)
            r[0][col] |= 0x80
            packed = np.uint64((col << 32) | (a_ent << 1) | (1 if nd == max(nd, na) else 0))
            inv = ~packed
            if inv > r[1]:
                r[1] = inv
    keys = sorted(rows)
    out_keys = np.array([(a << 32) | b for a, b in keys], dtype=np.uint64)
    words = np.zeros((len(keys), nb4 + 2), dtype=np.uint32)
    for i, k in enumerate(keys):
        words[i, :nb4] = rows[k][0].view(np.uint32)
        words[i, nb4:] = np.array([rows[k][1]], dtype=np.uint64).view(np.uint32)
    return out_keys, words, node_entry, pt.all_ids


def _reference(u):
    from oracle import hbond_port
    return wire_lists(hbond_port.run_wires(u, "protein", max_water=MAX_WATER, **KW).wire_lengths)


def test_decode_rows_single_process():
    u = _system().universe(N_FRAMES)
    keys, words, node_entry, ids = _device_rows(u, 0, N_FRAMES)
    cols, fd = split_wire_rows(words, N_FRAMES)
    got = decode_wire_rows(keys, cols, fd, node_entry, ids, N_FRAMES)
    ref = _reference(u)
    assert len(ref) > 30
    assert_same_wires(wire_lists(got), ref)


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
        u = _system().universe(N_FRAMES)
        lo, hi = shard_range(N_FRAMES, rank, world)
        keys, words, node_entry, ids = _device_rows(u, lo, hi)
        cols, fd = split_wire_rows(words, hi - lo, lo)
        keys, cols, fd = merge_wire_shards(keys, cols, fd, lo, hi, N_FRAMES)
        got = decode_wire_rows(keys, cols, fd, node_entry, ids, N_FRAMES)
        with open(os.path.join(out_dir, "rank%d.pkl" % rank), "wb") as f:
            pickle.dump(wire_lists(got), f)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_wires_equal_single_process(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref = _reference(_system().universe(N_FRAMES))
    for r in range(world):
        with open(os.path.join(str(tmp_path), "rank%d.pkl" % r), "rb") as f:
            assert_same_wires(pickle.load(f), ref, "rank %d of %d" % (r, world))


def test_wire_analysis_dump_and_restore(tmp_path):
    """basic.py:71-86 as cgraphs uses it for wires (proteingraphanalyser.py:499-509): results survive a dump / restore
    (no GPU needed: the results are planted)."""
    from cgraphs_b200.mdhbond import WireAnalysis
    u = _system().universe(3)
    wa = WireAnalysis("protein", u)
    wa.wire_lengths = {"PA-SER-1:PA-THR-2": np.array([1.0, np.inf, 0.0])}
    wa._set_results({k: v != np.inf for k, v in wa.wire_lengths.items()})
    f = str(tmp_path / "wires.pickle")
    wa.dump_to_file(f)
    wb = WireAnalysis(restore_filename=f)
    assert wb.nb_frames == 3 and list(wb.wire_lengths) == list(wa.wire_lengths)
    assert np.array_equal(wb.wire_lengths["PA-SER-1:PA-THR-2"], wa.wire_lengths["PA-SER-1:PA-THR-2"])
    assert wb.compute_average_water_per_wire() == {"PA-SER-1:PA-THR-2": 0.5}
    assert wb.filtered_graph.number_of_edges() == 1


def test_hbond_analysis_restore_keeps_derived_tables(tmp_path):
    """The reference pickles heavy2hydrogen / _all_ids / ... (basic.py:71-81); a restored object keeps them, keeps the
    constructor's runtime keyword arguments, and says clearly why it cannot run without its Universe."""
    from cgraphs_b200.mdhbond import HbondAnalysis
    u = _system().universe(3)
    a = HbondAnalysis("protein", u, device=0)
    a.initial_results = {"PA-SER-1:PA-THR-2": np.array([True, False, True])}
    f = str(tmp_path / "hba.pickle")
    a.dump_to_file(f)
    b = HbondAnalysis(restore_filename=f, device=3, pbc="ortho")
    assert b.heavy2hydrogen == a.heavy2hydrogen and b._all_ids == a._all_ids
    assert b._device == 3 and b._pbc == "ortho" and b._universe is None
    c = HbondAnalysis(restore_filename=f)
    assert c._device == 0 and c._pbc is False
    with pytest.raises(AssertionError, match="restored from a file"):
        b.set_hbonds_in_selection()
    assert a._topology is not None and a._universe is u            # dumping does not strip the live object
