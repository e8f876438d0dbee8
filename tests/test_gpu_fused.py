"""The one-CTA-per-frame kernel (hb_fused.cuh) and the general multi-kernel path (hb_general.cuh) must
give the same bits as the CPU oracle; frames that do not fit in shared memory must fall back without
losing or duplicating anything."""
import itertools

import numpy as np
import pytest

from cgraphs_b200 import synth
from cgraphs_b200.universe import Universe
from tests.helpers import assert_same_results

pytestmark = pytest.mark.gpu


def _run(u, selection, method, options, around=3.5, **kw):
    from cgraphs_b200.mdhbond import HbondAnalysis
    hba = HbondAnalysis(selection, u, native_options=options, **kw)
    if method == "in_selection":
        hba.set_hbonds_in_selection()
    else:
        hba.set_hbonds_in_selection_and_water_around(around)
    return hba


def _port(u, selection, method, around=3.5, **kw):
    from oracle import hbond_port
    return hbond_port.run(u, selection, method, around=around, **kw)


@pytest.mark.parametrize("seed", [0, 1])
def test_fused_and_general_match_oracle(seed):
    s = synth.make_system(5000, 40, 900 + seed, hydrogens=True, water="TIP3", n_chains=2)
    u = s.universe(9)
    for sel, method, ca in itertools.product(["protein", "protein or resname TIP3"],
                                             ["in_selection", "water_around"], [True, False]):
        kw = dict(check_angle=ca, additional_donors=['N'], additional_acceptors=['O'])
        ref = _port(u, sel, method, **kw).initial_results
        f = _run(u, sel, method, {"fused": 1}, **kw)
        g = _run(u, sel, method, {"fused": 0}, **kw)
        assert_same_results(f.initial_results, ref, "fused %s %s %s" % (sel, method, ca))
        assert_same_results(g.initial_results, ref, "general %s %s %s" % (sel, method, ca))
        assert g.last_run_stats["fused_frames"] == 0
        assert f.last_run_stats["fused_frames"] == 9, f.last_run_stats
        if f.last_run_stats["fused_fallbacks"] == 0:      # (a capacity retry launches the fused kernel again)
            assert f.last_run_stats["n_launches"] < g.last_run_stats["n_launches"]


@pytest.mark.parametrize("cut", [180.0, 179.0, 120.0])
def test_wide_angle_cutoffs_and_self_pairs(cut):
    """cut_angle = 180 admits cos_arg = -1: entries of one atom then pair with themselves (distance 0)."""
    s = synth.make_system(3000, 30, 55, hydrogens=True, water="TIP3")
    u = s.universe(3)
    for sel, method in [("protein", "in_selection"), ("protein or resname TIP3", "water_around")]:
        kw = dict(cut_angle=cut, additional_donors=['N'], additional_acceptors=['O'])
        ref = _port(u, sel, method, **kw).initial_results
        for fused in (1, 0):
            got = _run(u, sel, method, {"fused": fused}, **kw).initial_results
            assert_same_results(got, ref, "cut %s %s %s fused=%d" % (cut, sel, method, fused))


def test_fused_c3_half_scale():
    s, cfg = synth.config("C3", scale=0.5)
    u = s.universe(6)
    kw = dict(check_angle=True)
    ref = _port(u, "protein", "water_around", **kw).initial_results
    f = _run(u, "protein", "water_around", {"fused": 1}, **kw)
    assert_same_results(f.initial_results, ref)
    assert f.last_run_stats["fused_frames"] == 6 and f.last_run_stats["fused_fallbacks"] == 0
    assert f.last_run_stats["launches"]["frame_fused"] >= 1


def test_fused_falls_back_when_local_waters_do_not_fit():
    s = synth.make_system(6000, 50, 77)
    u = s.universe(5)
    ref = _port(u, "protein", "water_around").initial_results
    f = _run(u, "protein", "water_around", {"fused": 1, "fused_lw_cap": 8})
    assert_same_results(f.initial_results, ref)
    assert f.last_run_stats["fused_fallbacks"] >= 1


def test_fused_small_cell_budget_coarsens_grid():
    s = synth.make_system(6000, 50, 78)
    u = s.universe(3)
    ref = _port(u, "protein", "water_around").initial_results
    f = _run(u, "protein", "water_around", {"fused": 1, "fused_cell_cap": 64})
    assert_same_results(f.initial_results, ref)
    assert f.last_run_stats["fused_frames"] == 3


def test_fused_waters_without_regular_stride():
    """HOH oxygens only (stride 1) and a system whose waters are interleaved irregularly."""
    s = synth.make_system(3000, 60, 5, hydrogens=False, water="HOH")
    u = s.universe(1, static=True)
    kw = dict(check_angle=False, add_donors_without_hydrogen=True)
    ref = _port(u, "protein", "water_around", **kw).initial_results
    f = _run(u, "protein", "water_around", {"fused": 1}, **kw)
    assert_same_results(f.initial_results, ref)
    # irregular: drop every 7th water oxygen from the water group by renaming it
    topo = s.topology
    names = list(topo.names)
    wat = [i for i, (n, r) in enumerate(zip(topo.names, topo.resnames)) if r == "HOH"]
    for i in wat[::7]:
        names[i] = "OX"
    from cgraphs_b200.universe import Topology
    t2 = Topology(names, list(topo.resnames), list(topo.resids), list(topo.segids))
    u2 = Universe(t2, s.frames(0, 1, static=True))
    ref2 = _port(u2, "protein", "water_around", **kw).initial_results
    f2 = _run(u2, "protein", "water_around", {"fused": 1}, **kw)
    assert_same_results(f2.initial_results, ref2)
    assert f2.last_run_stats["fused_frames"] == 1


def test_fused_structure_batch():
    from cgraphs_b200.mdhbond import HbondBatch
    systems = synth.c5_structures(10, seed0=5100)
    unis = [s.universe(1, static=True) for s in systems]
    kw = dict(check_angle=False, add_donors_without_hydrogen=True)
    out = {}
    for fused in (1, 0):
        batch = HbondBatch(unis, "protein", native_options={"fused": fused}, **kw)
        batch.set_hbonds_in_selection_and_water_around(3.5)
        out[fused] = (batch.bond_keys, batch.occupancy.copy(), batch.last_run_stats)
    assert out[1][0] == out[0][0]
    assert np.array_equal(out[1][1], out[0][1])
    assert out[1][2]["fused_frames"] == 10 and out[0][2]["fused_frames"] == 0
    for u, col in zip(unis, range(10)):
        ref = _port(u, "protein", "water_around", **kw).initial_results
        got = {k for k, row in zip(out[1][0], out[1][1]) if row[col]}
        assert got == set(ref)


def test_fused_device_resident_unaligned_frames():
    """Frames at 12-byte granularity in device memory: slabs that cannot be 16-byte aligned inside the
    buffer use plain loads; results must not change with the buffer offset."""
    import torch
    from cgraphs_b200.mdhbond import HbondAnalysis
    s = synth.make_system(4001, 30, 41)          # odd atom count: frame starts rotate through all alignments
    nf = 12
    xyz = s.frames(0, nf)
    u = Universe(s.topology, xyz)
    hba = HbondAnalysis("protein", u, native_options={"fused": 1})
    hba.set_hbonds_in_selection_and_water_around(3.5)
    ctx = hba._context()
    ctx.begin(nf)
    ctx.submit(xyz)
    k0, w0 = ctx.results()
    for pad in (0, 1, 2, 3):
        buf = torch.zeros(xyz.size + pad, dtype=torch.float32, device="cuda")
        buf[pad:] = torch.from_numpy(xyz.reshape(-1)).cuda()
        torch.cuda.synchronize()
        ctx.begin(nf)
        ctx.submit_device(buf.data_ptr() + 4 * pad, xyz.shape[1], nf)
        k, w = ctx.results()
        assert np.array_equal(k, k0) and np.array_equal(w, w0), pad
    assert ctx.timers()["fused_frames"] == nf


def test_finalize_sort_paths_agree():
    """hb_finalize sorts small key sets in one CTA (bitonic) and large ones with the radix passes."""
    s, cfg = synth.config("C2", scale=0.6)
    u = s.universe(4)
    small = _run(u, "protein", "water_around", {})
    small_r = _run(u, "protein", "water_around", {"force_radix_sort": 1})
    assert 0 < len(small.initial_results) < 4096
    assert_same_results(small.initial_results, small_r.initial_results)
    big = _run(u, "protein or resname TIP3", "water_around", {})
    assert len(big.initial_results) > 4096
    assert_same_results(big.initial_results, _port(u, "protein or resname TIP3", "water_around").initial_results)


def test_h2d_skips_inert_atom_ranges():
    """Trajectory frames are copied to the device without the atom ranges no kernel reads (membrane filler)."""
    from cgraphs_b200.mdhbond import HbondAnalysis, _native
    s, cfg = synth.config("C3", scale=0.5)
    nf = 10
    xyz = s.frames(0, nf)
    u = Universe(s.topology, xyz)
    ref = _port(u, "protein", "water_around").initial_results
    full = _run(u, "protein", "water_around", {"copy_ranges": 0})
    part = _run(u, "protein", "water_around", {"copy_ranges": 1})
    assert_same_results(full.initial_results, ref)
    assert_same_results(part.initial_results, ref)
    assert full.last_run_stats["h2d_bytes"] == xyz.nbytes
    assert part.last_run_stats["h2d_bytes"] < 0.9 * xyz.nbytes
    # pinned source memory takes the direct (unstaged) copy path
    ctx = part._context()
    pin = _native.PinnedBuffer(xyz.shape)
    pin.array[...] = xyz
    ctx.begin(nf)
    ctx.submit_raw(pin.ptr, xyz.shape[1], nf)
    k1, w1 = ctx.results()
    ctx.begin(nf)
    ctx.submit(xyz)
    k2, w2 = ctx.results()
    pin.free()
    assert np.array_equal(k1, k2) and np.array_equal(w1, w2) and len(k1) == len(ref)


def test_merge_keys_kernel():
    """hb_merge_keys_device: merge step of the multi-GPU key exchange (emulated gather buffers on one GPU)."""
    import torch
    from cgraphs_b200.mdhbond import _native
    ctx = _native.Context(0)
    rng = np.random.default_rng(3)
    for world, cap, counts in [(3, 8, [5, 8, 0]), (8, 2048, [1507, 1490, 2048, 0, 1, 777, 1507, 1200]), (2, 4, [9, 2])]:
        gathered = rng.integers(0, 1 << 62, size=(world, cap + 1), dtype=np.int64)     # padding = garbage
        expect = []
        for w, c in enumerate(counts):
            keys = np.sort(rng.choice(5000, size=min(c, cap), replace=False).astype(np.int64) * 977)
            gathered[w, 0] = c
            gathered[w, 1:1 + len(keys)] = keys
            expect.append(keys)
        expect = np.unique(np.concatenate(expect))
        g = torch.from_numpy(gathered).cuda()
        out = torch.empty(world * cap, dtype=torch.int64, device="cuda")
        info = torch.empty(2, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        ctx.merge_keys_device(g.data_ptr(), world, cap, out.data_ptr(), info.data_ptr())
        ctx.sync()
        n_out, over = info.tolist()
        assert over == int(max(counts) > cap)
        assert n_out == len(expect)
        assert np.array_equal(out[:n_out].cpu().numpy(), expect)
    ctx.close()


def test_finalize_async_and_pack_block():
    """hb_finalize_async queues the finalisation; hb_pack_keys_device writes [count, keys...] from it; a bond table
    that overflows under a queued finalisation is redone (finalize_retries) with the same final result."""
    import torch
    from cgraphs_b200.mdhbond import HbondAnalysis
    s, cfg = synth.config("C2", scale=0.4)
    nf = 6
    xyz = s.frames(0, nf)
    u = Universe(s.topology, xyz)
    hba = HbondAnalysis("protein", u)
    hba.set_hbonds_in_selection_and_water_around(3.5)
    ctx = hba._context()
    ctx.begin(nf)
    ctx.submit(xyz)
    k0, w0 = ctx.results()
    cap = 256
    block = torch.full((cap + 1,), 7, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.begin(nf)
    ctx.submit(xyz)
    ctx.finalize_async()
    ctx.pack_keys_device(block.data_ptr(), cap)
    k1, w1 = ctx.results()
    assert np.array_equal(k0, k1) and np.array_equal(w0, w1)
    b = block.cpu().numpy()
    assert b[0] == len(k0) and np.array_equal(b[1:1 + len(k0)].view(np.uint64), k0) and (b[1 + len(k0):] == -1).all()
    # tiny table: the queued finalisation sees an overflow and is redone
    big = HbondAnalysis("protein or resname TIP3", u, native_options={"table_capacity": 1024})
    ref = _port(u, "protein or resname TIP3", "water_around").initial_results
    c2 = big._context()
    big.set_hbonds_in_selection_and_water_around(3.5)
    assert_same_results(big.initial_results, ref)
    assert big.last_run_stats["table_grows"] >= 1


def test_general_path_pair_buffer_overflow_and_fallback():
    """The general path stages point pairs in a batch-wide buffer; a buffer that is too small halves the sub-batch
    and finally falls back to the per-point pair kernel - same bits every way."""
    s, cfg = synth.config("C2", scale=0.3)
    u = s.universe(24)
    sel = "protein or resname TIP3"
    ref = _port(u, sel, "water_around").initial_results
    for opts in ({"fused": 0}, {"fused": 0, "frame_pairs": 0}, {"fused": 0, "frame_pairs": 0, "pair_buf_bytes": 0},
                 {"fused": 0, "frame_pairs": 0, "pair_buf_bytes": 1 << 16}, {"fused": 0, "frame_pairs": 0, "pair_buf_bytes": 4096}):
        got = _run(u, sel, "water_around", opts)
        assert_same_results(got.initial_results, ref, str(opts))
        assert got.last_run_stats["fused_frames"] == 0


@pytest.mark.parametrize("distance,around,cut", [(2.8, 2.0, 20.0), (3.0, 2.8, 30.0), (3.5, 6.0, 45.0), (4.2, 3.5, 90.0),
                                                 (3.0, 9.0, 60.0)])
def test_other_cutoffs_and_angles(distance, around, cut):
    """Cut-offs other than cgraphs' defaults: around_radius above and below distance, narrow and wide angles."""
    s = synth.make_system(4500, 35, 400 + int(distance * 10))
    u = s.universe(4)
    for sel in ("protein", "protein or resname TIP3"):
        kw = dict(distance=distance, cut_angle=cut)
        ref = _port(u, sel, "water_around", around=around, **kw).initial_results
        assert len(ref) > 0 or sel == "protein"         # (the narrowest case leaves the protein without bonds)
        for fused in (1, 0):
            got = _run(u, sel, "water_around", {"fused": fused}, around=around, **kw)
            assert_same_results(got.initial_results, ref, "d=%s a=%s cut=%s %s fused=%d" % (distance, around, cut, sel, fused))


@pytest.mark.parametrize("n_atoms", [16000, 60000])
def test_frame_pairs_capacity_levels(n_atoms):
    """General path, pair search as one CTA per frame (k_frame_pairs): frames with more points than two CTAs per SM
    can hold move to one CTA per SM, larger ones to the multi-kernel pair search - same bits every way."""
    s = synth.make_system(n_atoms, 50, 1234)
    u = s.universe(3)
    sel = "protein or resname TIP3"
    ref = _port(u, sel, "water_around").initial_results
    assert len(ref) > 1000
    for opts in ({"fused": 0}, {"fused": 0, "frame_pairs": 0}):
        got = _run(u, sel, "water_around", opts)
        assert_same_results(got.initial_results, ref, str(opts))



def test_fused_full_scale_passes_are_reproducible():
    """Three device-resident passes over 2,048 full-scale C3 frames (four resident CTAs per SM, every stage of the tile
    ring refilled ~30 times per frame) and one pass from host memory give the same keys and occupancy rows: a refill of a
    ring stage that overtakes a lane's reads of it shows up as a few frames that differ from run to run."""
    import torch
    from cgraphs_b200.mdhbond import HbondAnalysis, _native
    from cgraphs_b200.mdhbond.calib import cos_threshold
    from cgraphs_b200.mdhbond.topology import pack_systems
    system, cfg = synth.config("C3")
    nf, na = 2048, system.n_atoms
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
    runs = []
    for _ in range(3):
        ctx.begin(nf)
        ctx.submit_device(d_xyz.data_ptr(), na, nf)
        ctx.finalize()
        assert ctx.timers()["fused_frames"] == nf
        runs.append(ctx.results())
    host = d_xyz.cpu().numpy()
    ctx.begin(nf)
    ctx.submit_raw(host.ctypes.data, na, nf)
    ctx.finalize()
    runs.append(ctx.results())
    k0, w0 = runs[0]
    assert len(k0) > 1000
    for k, w in runs[1:]:
        assert np.array_equal(k, k0) and np.array_equal(w, w0)
