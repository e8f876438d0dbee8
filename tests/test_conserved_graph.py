"""ConservedGraph.get_conserved_graph (conservedgraph.py:114-170, graph_type "hbond") including the
renaming of waters that sit on a conserved-water cluster centre (conservedgraph.py:131-149).

CPU: the restatement oracle.hbond_port.conserved_graph against the LIVE reference function (run on prepared objects
through oracle/ref_harness.py), and the host logic of HbondBatch.get_conserved_graph on planted matrices against both.
GPU: HbondBatch end to end against the restatement."""
import numpy as np
import pytest

from cgraphs_b200 import synth
from cgraphs_b200.universe import Universe
from oracle import hbond_port, ref_harness

KW = dict(check_angle=False, add_donors_without_hydrogen=True)


def _structures(n=9, seed=901):
    s = synth.make_system(3000, 60, seed, hydrogens=False, water="HOH")
    xyz = s.frames(0, n)
    return s, [Universe(s.topology, xyz[i:i + 1]) for i in range(n)]


def _water_positions(u):
    t = u.topology
    pos = np.asarray(u.trajectory.frames.frame(0))
    return {int(t.resids[i]): pos[i] for i in range(t.n_atoms) if t.resnames[i] == "HOH" and t.names[i] == "O"}


def _centres(unis, k=6, seed=3):
    """Cluster centres: positions of a few waters of the first structure that are H-bonded to the protein there."""
    g = hbond_port.run(unis[0], "protein", "water_around", **KW).filtered_graph
    wat = sorted({n for e in g.edges for n in e if n.split('-')[1] == "HOH"})
    rng = np.random.default_rng(seed)
    pick = rng.choice(len(wat), size=min(k, len(wat)), replace=False)
    pos = _water_positions(unis[0])
    return {"X-w-%d" % j: tuple(float(x) for x in pos[int(wat[i].split('-')[2])]) for j, i in enumerate(pick)}


def _canon(edges):
    out = set()
    for e in edges:
        e = list(e)
        out.add(frozenset(e) if len(set(e)) > 1 else (e[0],))
    return out


@pytest.mark.reference
@pytest.mark.parametrize("cth,eps", [(0.9, 1.5), (0.5, 1.5), (0.2, 0.8), (0.5, 3.0)])
def test_port_conserved_graph_equals_reference(cth, eps):
    import warnings
    warnings.filterwarnings("ignore")
    s, unis = _structures()
    graphs = [hbond_port.run(u, "protein", "water_around", **KW).filtered_graph for u in unis]
    for centres in (None, _centres(unis)):
        objs = {"s%d" % i: {"graph": g, "structure": u} for i, (g, u) in enumerate(zip(graphs, unis))}
        rn, re = ref_harness.reference_conserved_graph(objs, centres or {}, cth, eps)
        pn, pe = hbond_port.conserved_graph(graphs, [_water_positions(u) for u in unis], centres, cth, eps)
        assert set(str(x) for x in rn) == pn
        assert _canon(re) == pe
        assert len(pe) > 0
    if eps >= 1.5:
        assert any("X-w-" in x for e in pe for x in e)           # the renaming really happened


def test_batch_host_logic_on_planted_matrix():
    """HbondBatch.get_conserved_graph on a planted bond x structure matrix (no GPU): equal to the restatement."""
    from cgraphs_b200.mdhbond import HbondBatch
    from cgraphs_b200.mdhbond.batch import StructureResult
    s, unis = _structures(7)
    runs = [hbond_port.run(u, "protein", "water_around", **KW) for u in unis]
    keys = sorted({k for r in runs for k in r.initial_results})
    occ = np.array([[k in r.initial_results for r in runs] for k in keys], dtype=bool)
    batch = HbondBatch(unis, "protein", **KW)
    batch.bond_keys, batch.occupancy = keys, occ
    batch.results = [StructureResult(r.initial_results, True) for r in runs]
    centres = _centres(unis)
    for cth, eps, cc in [(0.9, 1.5, None), (0.5, 1.5, centres), (0.3, 3.0, centres), (0.9, 0.5, centres)]:
        nodes, edges = batch.get_conserved_graph(cth, eps=eps, reference_coordinates=cc)
        pn, pe = hbond_port.conserved_graph([r.filtered_graph for r in runs], [_water_positions(u) for u in unis], cc, cth, eps)
        assert set(nodes) == pn and _canon(edges) == pe, (cth, eps, cc is not None)


@pytest.mark.gpu
def test_gpu_batch_conserved_graph_with_water_clusters():
    from cgraphs_b200.mdhbond import HbondBatch
    s, unis = _structures(9)
    batch = HbondBatch(unis, "protein", **KW)
    batch.set_hbonds_in_selection_and_water_around(3.5)
    graphs = [hbond_port.run(u, "protein", "water_around", **KW).filtered_graph for u in unis]
    wp = [_water_positions(u) for u in unis]
    centres = _centres(unis)
    for cth, eps, cc in [(0.9, 1.5, None), (0.5, 1.5, centres), (0.2, 3.0, centres)]:
        nodes, edges = batch.get_conserved_graph(cth, eps=eps, reference_coordinates=cc)
        pn, pe = hbond_port.conserved_graph(graphs, wp, cc, cth, eps)
        assert set(nodes) == pn and _canon(edges) == pe
        assert len(pe) > 0
