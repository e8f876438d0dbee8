"""XTC frame sources (basic.py:46-47 reads .xtc through MDAnalysis = xdrfile + float32 unit conversion).

CPU tests: the oracle restatement (oracle/xtc_codec.c) against itself on every adaptive path of the format, the
product's host writer / reader / index (libhbgpu.so, no GPU needed) bit-exact against the oracle.
GPU tests: the device decoder (k_xtc_decode through hb_submit_frames_xtc) bit-exact against the oracle, and
HbondAnalysis on .pdb + .xtc files equal to the oracle H-bond port fed with the oracle-decoded coordinates.
"""
import numpy as np
import pytest

from cgraphs_b200 import synth
from cgraphs_b200.universe import Universe
from tests.helpers import assert_same_results


def _quantised(nm, precision):
    p = np.float32(precision)
    lf = np.where(nm >= 0, nm * p + np.float32(0.5), nm * p - np.float32(0.5)).astype(np.float32)
    return lf.astype(np.int32)


def _expected(nm, precision, scale):
    q = _quantised(nm, precision).astype(np.float32)
    return (q * np.float32(1.0 / precision)) * np.float32(scale)


def _cases():
    """name -> (frames in nm float32 [F, N, 3], precision): every branch of the codec."""
    rng = np.random.default_rng(7)
    out = {}
    s = synth.make_system(3000, 30, 42)
    out["solvated"] = ((s.frames(0, 3) * np.float32(0.1)).astype(np.float32), 1000.0)        # water runs, smallidx moves
    out["gas"] = (rng.uniform(0, 30, size=(2, 500, 3)).astype(np.float32), 1000.0)            # no small runs at all
    out["negative"] = (rng.normal(0, 3, size=(2, 700, 3)).astype(np.float32), 1000.0)         # negative coordinates
    out["tiny"] = (rng.uniform(0, 5, size=(3, 9, 3)).astype(np.float32), 1000.0)              # <= 9 atoms: plain floats
    out["ten"] = (rng.uniform(0, 5, size=(2, 10, 3)).astype(np.float32), 1000.0)              # smallest compressed frame
    chain = np.cumsum(rng.normal(0, 0.05, size=(1, 4000, 3)), axis=1).astype(np.float32)       # one long run of small steps
    out["chain"] = (chain, 1000.0)
    cl = (rng.uniform(0, 4, size=(2, 60, 1, 3)) + rng.normal(0, 0.004, size=(2, 60, 20, 3))).reshape(2, 1200, 3)
    out["clusters"] = (cl.astype(np.float32), 1000.0)                                         # runs of the maximum length 8
    wide = rng.uniform(-9000, 9000, size=(2, 300, 3)).astype(np.float32)                       # range > 2^24: three plain ints
    out["wide"] = (wide, 1000.0)
    big = rng.uniform(0, 16000, size=(2, 300, 3)).astype(np.float32)                           # 3 x 24 bits: 72-bit triples
    out["bits72"] = (big, 1000.0)
    out["coarse"] = ((s.frames(0, 2) * np.float32(0.1)).astype(np.float32), 100.0)            # other precision
    out["fine"] = ((s.frames(0, 2) * np.float32(0.1)).astype(np.float32), 100000.0)           # big integers, > 32-bit small triples
    ident = np.broadcast_to(rng.uniform(0, 3, size=(1, 1, 3)), (1, 64, 3)).astype(np.float32)  # all atoms on one point
    out["identical"] = (np.ascontiguousarray(ident), 1000.0)
    return out


CASES = _cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_roundtrip_and_product_host_codec(name):
    """decode(encode(x)) is the quantised x exactly; the product's writer emits the same bytes as the oracle's; the
    product's host reader and index agree with the oracle."""
    from oracle import xtc_port
    from cgraphs_b200.mdhbond import _native
    nm, precision = CASES[name]
    buf, off = xtc_port.encode(nm, precision=precision, box_nm=np.diag([5.0, 6.0, 7.0]))
    na = nm.shape[1]
    off2, na2 = xtc_port.index(buf)
    assert np.array_equal(off, off2) and na2 == na
    for scale in (1.0, 10.0):
        dec, box = xtc_port.decode(buf, off, na, scale=scale)
        want = _expected(nm, precision, scale) if na > 9 else nm * np.float32(scale)
        assert np.array_equal(dec, want), name
        assert np.array_equal(box[0], np.diag([5.0, 6.0, 7.0]).astype(np.float32))
    assert np.abs(xtc_port.decode(buf, off, na, 1.0)[0] - nm).max() <= (0.5 / precision) * 1.001 + 4 * np.spacing(np.abs(nm).max())
    # product host side
    pbuf, poff = _native.xtc_encode(nm, in_scale=1.0, precision=precision, box_nm=np.diag([5.0, 6.0, 7.0]))
    assert np.array_equal(pbuf, buf) and np.array_equal(poff, off), "writer bytes differ from the oracle's (%s)" % name
    ioff, ina = _native.xtc_index(buf)
    assert np.array_equal(ioff, off) and ina == na
    dec, _ = xtc_port.decode(buf, off, na, scale=10.0)
    for f in range(len(nm)):
        x, b = _native.xtc_decode_host(buf, off[f], na, 10.0)
        assert np.array_equal(x, dec[f]), name
        assert np.array_equal(b, np.diag([5.0, 6.0, 7.0]).astype(np.float32))


def test_format_landmarks():
    """What a reader of real files relies on: XDR big-endian header, magic 1995, record length = 92 + padded bit stream."""
    from oracle import xtc_port
    nm, precision = CASES["solvated"]
    buf, off = xtc_port.encode(nm[:1], precision=precision, first_step=7, dt=2.0)
    be = lambda o: int.from_bytes(bytes(buf[o:o + 4]), "big", signed=True)
    assert be(0) == 1995 and be(4) == nm.shape[1] and be(8) == 7 and be(52) == nm.shape[1]
    assert np.frombuffer(bytes(buf[12:16]), ">f4")[0] == 14.0 and np.frombuffer(bytes(buf[56:60]), ">f4")[0] == 1000.0
    q = _quantised(nm[0], precision)
    assert [be(60 + 4 * k) for k in range(3)] == list(q.min(axis=0)) and [be(72 + 4 * k) for k in range(3)] == list(q.max(axis=0))
    assert 9 <= be(84) <= 72 and len(buf) == 92 + ((be(88) + 3) & ~3) and len(buf) % 4 == 0
    assert len(buf) < 0.4 * nm[0].nbytes                          # compresses a solvated system about 3x


def test_malformed_records_are_rejected():
    from oracle import xtc_port
    from cgraphs_b200.mdhbond import _native
    nm, precision = CASES["solvated"]
    buf, off = xtc_port.encode(nm[:2], precision=precision)
    bad = buf.copy()
    bad[3] ^= 1                                                   # magic
    assert len(_native.xtc_index(bad)[0]) == 1                    # no frames
    o, na = _native.xtc_index(buf[:off[2] - 8])                   # truncated second record: one complete frame
    assert len(o) == 2 and na == nm.shape[1]
    with pytest.raises(_native.HbError):
        _native.xtc_decode_host(buf, off[0], nm.shape[1] + 1, 10.0)


def test_xtc_files_as_frame_source(tmp_path):
    """write_xtc / XTCFrames / universe_from_files: frames in Angstrom equal the oracle's decode; several files chain."""
    from oracle import xtc_port
    from cgraphs_b200 import io
    s = synth.make_system(2500, 24, 5, dimensions_kind="tric")
    xyz = s.frames(0, 7)
    pdb, x1, x2 = str(tmp_path / "s.pdb"), str(tmp_path / "a.xtc"), str(tmp_path / "b.xtc")
    io.write_pdb(pdb, s.topology, xyz[0], s.dimensions)
    io.write_xtc(x1, xyz[:4], s.dimensions)
    io.write_xtc(x2, xyz[4:], s.dimensions)
    u = io.universe_from_files(pdb, [x1, x2])
    fr = u.trajectory.frames
    assert len(fr) == 7 and fr.n_atoms == 2500
    raw = np.fromfile(x1, dtype=np.uint8)
    off, na = xtc_port.index(raw)
    want, box = xtc_port.decode(raw, off, na, scale=10.0)
    assert np.array_equal(fr.block(0, 4), want) and np.array_equal(fr.frame(5), u.trajectory.frames.parts[1].frame(1))
    assert np.abs(fr.block(0, 7) - xyz).max() <= 0.0051
    d = fr.dimensions_of(0)
    assert np.allclose(d, s.dimensions, atol=2e-3)
    r, o = fr.xtc_block(1, 4)
    assert fr.xtc_count(2, 7) == 2 and len(o) == 4 and o[0] == 0 and o[-1] == len(r)
    u1 = io.universe_from_files(pdb, x1)
    assert len(u1.trajectory) == 4


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_decode_bit_exact(name):
    """k_xtc_decode writes exactly the floats the oracle decodes (pageable and pinned sources, with and without
    atom-range filtering): checked through a presence run whose device input buffer is read back."""
    import torch
    from oracle import xtc_port
    from cgraphs_b200.mdhbond import _native
    nm, precision = CASES[name]
    nf, na = nm.shape[:2]
    buf, off = xtc_port.encode(nm, precision=precision)
    want, _ = xtc_port.decode(buf, off, na, scale=10.0)
    got = _native.xtc_decode_device(buf, off, na, scale=10.0)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "%s: %d floats differ" % (
        name, int((got.view(np.uint32) != want.view(np.uint32)).sum()))


@pytest.mark.gpu
def test_gpu_hbonds_from_xtc_files(tmp_path):
    """HbondAnalysis('protein', 's.pdb', 't.xtc'): same result dict as the oracle port on the oracle-decoded frames,
    on both kernel paths, with several batches per run."""
    from oracle import hbond_port, xtc_port
    from cgraphs_b200 import io
    from cgraphs_b200.mdhbond import HbondAnalysis
    s, cfg = synth.config("C3", scale=0.2)
    nf = 24
    xyz = s.frames(0, nf)
    pdb, xtc = str(tmp_path / "s.pdb"), str(tmp_path / "t.xtc")
    io.write_xtc(xtc, xyz, s.dimensions)
    raw = np.fromfile(xtc, dtype=np.uint8)
    off, na = xtc_port.index(raw)
    dec, _ = xtc_port.decode(raw, off, na, scale=10.0)
    io.write_pdb(pdb, s.topology, dec[0], s.dimensions)
    u_ref = Universe(io.read_pdb(pdb)[0], dec, s.dimensions)
    ref = hbond_port.run(u_ref, "protein", "water_around", around=3.5).initial_results
    assert len(ref) > 50
    for opts in ({}, {"fused": 0}, {"max_batch_frames": 5, "copy_ranges": 0}):
        hba = HbondAnalysis("protein", pdb, xtc, native_options=opts)
        assert hba.nb_frames == nf
        hba.set_hbonds_in_selection_and_water_around(3.5)
        assert_same_results(hba.initial_results, ref, "xtc %s" % opts)
        assert hba.last_run_stats["h2d_bytes"] <= len(raw)
    hba = HbondAnalysis("protein", pdb, xtc, start=3, stop=20, pbc=True)
    hba.set_hbonds_in_selection_and_water_around(3.5)
    ref2 = hbond_port.run(Universe(u_ref.topology, dec[3:20], s.dimensions), "protein", "water_around", around=3.5).initial_results
    assert_same_results(hba.initial_results, ref2, "xtc slice + pbc on a skin system")


@pytest.mark.gpu
def test_gpu_malformed_xtc_fails_loudly():
    from oracle import xtc_port
    from cgraphs_b200.mdhbond import _native
    nm, precision = CASES["solvated"]
    buf, off = xtc_port.encode(nm, precision=precision)
    bad = buf.copy()
    bad[off[1] + 3] ^= 1                                          # magic of the second frame
    with pytest.raises(_native.HbError, match="malformed XTC"):
        _native.xtc_decode_device(bad, off, nm.shape[1], scale=10.0)
    bad = buf.copy()
    bad[off[1] + 84:off[1] + 88] = [0, 0, 0, 99]                  # smallidx out of range
    with pytest.raises(_native.HbError, match="malformed XTC"):
        _native.xtc_decode_device(bad, off, nm.shape[1], scale=10.0)
