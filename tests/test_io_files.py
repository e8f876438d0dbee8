"""PDB / DCD readers (cgraphs_b200/io.py): round trips on CPU; the planar DCD path through the kernels on the GPU."""
import numpy as np
import pytest

from cgraphs_b200 import synth
from cgraphs_b200.io import DCDFrames, read_pdb, universe_from_files, write_dcd, write_pdb
from tests.helpers import assert_same_results


@pytest.fixture(autouse=True)
def _no_mdanalysis(monkeypatch):
    """These tests exercise the built-in readers: hide MDAnalysis (absent in this image; the live-reference harness of
    other tests installs a stub module under that name)."""
    import sys
    monkeypatch.setitem(sys.modules, "MDAnalysis", None)


def _files(tmp_path, n_frames=7, kind="tric"):
    s = synth.make_system(3200, 25, 61, dimensions_kind=kind)
    xyz = s.frames(0, n_frames)
    pdb, dcd = str(tmp_path / "sys.pdb"), str(tmp_path / "traj.dcd")
    write_pdb(pdb, s.topology, xyz[0], s.dimensions)
    write_dcd(dcd, xyz, s.dimensions)
    return s, xyz, pdb, dcd


def test_pdb_dcd_round_trip(tmp_path):
    s, xyz, pdb, dcd = _files(tmp_path)
    topo, x0, dims = read_pdb(pdb)
    t = s.topology
    assert list(topo.names) == list(t.names) and list(topo.resnames) == list(t.resnames)
    assert list(topo.segids) == list(t.segids) and np.array_equal(topo.resids, t.resids)
    assert np.allclose(x0, xyz[0], atol=6e-4)                     # PDB keeps three decimals
    assert np.allclose(dims, s.dimensions, atol=1e-2)
    fr = DCDFrames(dcd)
    assert len(fr) == len(xyz) and fr.n_atoms == xyz.shape[1]
    assert np.array_equal(fr.block(0, len(xyz)), xyz)             # DCD keeps float32 exactly
    assert np.array_equal(fr.frame(3), xyz[3])
    assert np.allclose(fr.dimensions_of(2), s.dimensions, atol=1e-4)
    raw, stride, xo, yo, zo = fr.planar_block(2, 5)
    assert len(raw) == 3 * stride and xo % 4 == 0 and yo % 4 == 0 and zo % 4 == 0 and stride % 4 == 0
    assert np.array_equal(np.asarray(raw[stride + yo:stride + yo + 4 * fr.n_atoms]).view("<f4"), xyz[3][:, 1])
    u = universe_from_files(pdb, dcd)
    assert len(u.trajectory) == len(xyz) and len(u.select_atoms("protein")) == len(s.universe(1).select_atoms("protein"))


@pytest.mark.gpu
@pytest.mark.parametrize("pbc", [False, True])
def test_gpu_reads_dcd_records_directly(tmp_path, pbc):
    from cgraphs_b200.mdhbond import HbondAnalysis
    from oracle import hbond_port
    s, xyz, pdb, dcd = _files(tmp_path, n_frames=9)
    ref = hbond_port.run(s.universe(9), "protein", "water_around", around=3.5).initial_results
    for opts in ({"fused": 1}, {"fused": 0}, {"copy_ranges": 0}):
        hba = HbondAnalysis("protein", pdb, dcd, pbc=pbc, native_options=opts)       # file paths, no MDAnalysis
        hba.set_hbonds_in_selection_and_water_around(3.5)
        assert_same_results(hba.initial_results, ref, "dcd %s pbc=%s" % (opts, pbc))
        assert hba.nb_frames == 9
    sl = HbondAnalysis("protein", pdb, dcd, start=2, stop=8, step=1)
    sl.set_hbonds_in_selection_and_water_around(3.5)
    assert {k: v.tobytes() for k, v in sl.initial_results.items()} == {k: v[2:8].tobytes() for k, v in ref.items() if v[2:8].any()}


@pytest.mark.gpu
def test_gpu_wires_and_presence_from_files(tmp_path):
    """WireAnalysis / HydrationAnalysis constructed from a PDB + DCD pair like cgraphs does (file paths, frame records
    decoded on the device for the wires)."""
    from cgraphs_b200.mdhbond import HydrationAnalysis, WireAnalysis
    from oracle import hbond_port
    from tests.helpers import assert_same_wires, wire_lists
    s, xyz, pdb, dcd = _files(tmp_path, n_frames=9)
    u = s.universe(9)
    kw = dict(additional_donors=['N'], additional_acceptors=['O'])
    wa = WireAnalysis("protein", pdb, dcd, **kw)
    wa.set_water_wires(max_water=3)
    ref = hbond_port.run_wires(u, "protein", max_water=3, **kw)
    assert len(ref.wire_lengths) > 5
    assert_same_wires(wire_lists(wa.wire_lengths), wire_lists(ref.wire_lengths))
    ha = HydrationAnalysis("protein", pdb, dcd)
    ha.set_presence_around(4.0)
    refp = hbond_port.run_presence_around(u, "protein", 4.0)
    assert {k: v.tobytes() for k, v in ha.initial_results.items()} == {k: v.tobytes() for k, v in refp.items()}


def test_pdb_atom_types(tmp_path):
    """Element columns win, names with leading digits are hydrogens (the donor rule reads the type)."""
    p = tmp_path / "t.pdb"
    p.write_text(
        "ATOM      1  N   SER A   1       0.000   0.000   0.000  1.00  0.00           N\n"
        "ATOM      2 1HB  SER A   1       1.000   0.000   0.000  1.00  0.00\n"
        "ATOM      3  HG  SER A   1       0.000   1.000   0.000  1.00  0.00           H\n"
        "HETATM    4  O   HOH A   2       3.000   0.000   0.000  1.00  0.00           O\n"
        "END\n")
    topo, xyz, dims = read_pdb(str(p))
    assert list(topo.types) == ["N", "H", "H", "O"] and list(topo.names) == ["N", "1HB", "HG", "O"]
    assert dims is None and xyz.shape == (4, 3)


def test_psf_and_several_dcd_files(tmp_path):
    """What cgraphs' trajectory workflow passes: a PSF topology and a list of DCD files (proteingraphanalyser.py:472-476)."""
    from cgraphs_b200.io import ConcatDCDFrames, read_psf, write_psf
    s = synth.make_system(1800, 16, 7)
    xyz = s.frames(0, 11)
    psf = str(tmp_path / "sys.psf")
    write_psf(psf, s.topology)
    topo = read_psf(psf)
    t = s.topology
    assert list(topo.names) == list(t.names) and list(topo.resnames) == list(t.resnames)
    assert list(topo.segids) == list(t.segids) and np.array_equal(topo.resids, t.resids) and list(topo.types) == list(t.types)
    parts = [(0, 4), (4, 5), (5, 11)]
    dcds = []
    for k, (a, b) in enumerate(parts):
        p = str(tmp_path / ("part%d.dcd" % k))
        write_dcd(p, xyz[a:b], s.dimensions)
        dcds.append(p)
    fr = ConcatDCDFrames(dcds)
    assert len(fr) == 11 and np.array_equal(fr.block(0, 11), xyz) and np.array_equal(fr.frame(4), xyz[4])
    assert fr.planar_count(2, 9) == 2 and fr.planar_count(4, 9) == 1 and fr.planar_count(5, 9) == 4
    raw, stride, xo, yo, zo = fr.planar_block(5, 9)
    assert len(raw) == 4 * stride
    assert np.array_equal(np.asarray(raw[xo:xo + 4 * fr.n_atoms]).view("<f4"), xyz[5][:, 0])
    u = universe_from_files(psf, dcds)
    assert len(u.trajectory) == 11
    # the drop-in classes take the same arguments (constructor only: no GPU here) and see the same topology as the oracle
    from cgraphs_b200.mdhbond import WireAnalysis
    from oracle import hbond_port
    wa = WireAnalysis("protein", psf, dcds)
    assert wa.nb_frames == 11
    pt = hbond_port.PortTopology(s.universe(11), "protein")
    assert wa._nb_donors == pt.nD and wa._nb_acceptors == pt.nA and list(wa._all_ids) == list(pt.all_ids)
    with pytest.raises(ValueError):
        universe_from_files(psf)


@pytest.mark.gpu
def test_gpu_psf_and_several_dcd_files(tmp_path):
    """PSF + a list of DCD files through the planar path: blocks end at file boundaries."""
    from cgraphs_b200.io import write_psf
    from cgraphs_b200.mdhbond import HbondAnalysis, WireAnalysis
    from oracle import hbond_port
    from tests.helpers import assert_same_wires, wire_lists
    s = synth.make_system(3200, 25, 61, dimensions_kind="tric")
    xyz = s.frames(0, 11)
    psf = str(tmp_path / "sys.psf")
    write_psf(psf, s.topology)
    dcds = []
    for k, (a, b) in enumerate([(0, 4), (4, 5), (5, 11)]):
        p = str(tmp_path / ("part%d.dcd" % k))
        write_dcd(p, xyz[a:b], s.dimensions)
        dcds.append(p)
    u = s.universe(11)
    hba = HbondAnalysis("protein", psf, dcds)
    hba.set_hbonds_in_selection_and_water_around(3.5)
    assert_same_results(hba.initial_results, hbond_port.run(u, "protein", "water_around", around=3.5).initial_results)
    kw = dict(additional_donors=['N'], additional_acceptors=['O'])
    wa = WireAnalysis("protein", psf, dcds, **kw)
    wa.set_water_wires(max_water=3)
    assert_same_wires(wire_lists(wa.wire_lengths), wire_lists(hbond_port.run_wires(u, "protein", max_water=3, **kw).wire_lengths))
