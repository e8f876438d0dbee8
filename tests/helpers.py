"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np


def results_bytes(results):
    """Canonical comparable form of a result dict: key -> bytes of the bool vector."""
    return {k: np.asarray(v, dtype=bool).tobytes() for k, v in results.items()}


def assert_same_results(a, b, what=""):
    ra, rb = results_bytes(a), results_bytes(b)
    if ra == rb:
        return
    only_a = sorted(set(ra) - set(rb))
    only_b = sorted(set(rb) - set(ra))
    diff = sorted(k for k in set(ra) & set(rb) if ra[k] != rb[k])
    raise AssertionError("%s results differ: %d keys only in first %s, %d only in second %s, %d with different "
                         "frames %s" % (what, len(only_a), only_a[:5], len(only_b), only_b[:5], len(diff), diff[:5]))


def graph_edges(g):
    return {frozenset(e) for e in g.edges()}


def results_to_lists(results):
    """Compact JSON form: key -> list of frame indices where the bond exists."""
    return {k: np.nonzero(np.asarray(v, dtype=bool))[0].tolist() for k, v in sorted(results.items())}


def lists_to_results(d, nb_frames):
    out = {}
    for k, idx in d.items():
        row = np.zeros(nb_frames, dtype=bool)
        row[idx] = True
        out[k] = row
    return out


import json
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_FIXTURES = ["peptide_traj", "crystal_static", "kat_distance", "kat_angle"]


def load_golden(name):
    """(Universe, runs) of a golden fixture."""
    from cgraphs_b200.universe import Topology, Universe
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    topo = Topology([str(s) for s in z["names"]], [str(s) for s in z["resnames"]], z["resids"],
                    [str(s) for s in z["segids"]])
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        runs = json.load(f)["runs"]
    return Universe(topo, z["xyz"]), runs


def golden_cases():
    out = []
    for name in GOLDEN_FIXTURES:
        with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
            n = len(json.load(f)["runs"])
        out += [(name, i) for i in range(n)]
    return out


def ion_system(seed=3, n_frames=3):
    """A synthetic system with 40 ions (SOD / CLA) placed next to protein and water atoms."""
    from cgraphs_b200 import synth
    from cgraphs_b200.universe import Topology, Universe
    s = synth.make_system(3000, 25, seed, n_filler=60)
    t = s.topology
    names, resnames, resids, segids = list(t.names), list(t.resnames), list(t.resids), list(t.segids)
    fill = [i for i, r in enumerate(resnames) if r == "POPC"]
    for k, i in enumerate(fill[:40]):
        nm = "SOD" if k % 2 == 0 else "CLA"
        names[i] = nm
        resnames[i] = nm
        resids[i] = 5000 + k
        segids[i] = "ION"
    xyz = s.frames(0, n_frames)
    rng = np.random.default_rng(seed)
    targets = rng.choice(len(names) - len(fill), size=40, replace=False)
    for k, i in enumerate(fill[:40]):
        xyz[:, i, :] = xyz[:, targets[k], :] + rng.normal(0, 1.4, size=(n_frames, 3)).astype(np.float32)
    return Universe(Topology(names, resnames, resids, segids), xyz, s.dimensions)


def port_with_ions(u, selection, method, ions, around=3.5, **kw):
    """Oracle result of a set_hbonds_* call on an analysis constructed with ions= (hbond.py:135,289)."""
    from oracle import hbond_port
    res = dict(hbond_port.run(u, selection, method, around=around, **kw).initial_results)
    res.update(hbond_port.run_ions(u, selection, ions, include_water=(method == "water_around"), **kw))
    return res


def load_golden_wires(name="wires_traj"):
    """(Universe, runs) of a golden water-wire fixture (tests/golden/make_golden_wires.py)."""
    from cgraphs_b200.universe import Topology, Universe
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        doc = json.load(f)
    runs = doc["runs"]
    z = np.load(os.path.join(GOLDEN_DIR, doc.get("coordinates", name + ".npz")))     # (wires_hull shares wires_traj.npz)
    topo = Topology([str(s) for s in z["names"]], [str(s) for s in z["resnames"]], z["resids"],
                    [str(s) for s in z["segids"]])
    return Universe(topo, z["xyz"]), runs


def wire_lists(wire_lengths):
    """key -> list over frames, -1 for inf (the golden form)."""
    return {k: [(-1 if not np.isfinite(x) else int(x)) for x in v] for k, v in wire_lengths.items()}


def assert_same_wires(got, want, what=""):
    """Both: key -> per-frame list (-1 = none). Keys must agree in orientation."""
    if got == want:
        return
    only_a = sorted(set(got) - set(want))
    only_b = sorted(set(want) - set(got))
    diff = sorted(k for k in set(got) & set(want) if list(got[k]) != list(want[k]))
    detail = ["%s: got %s want %s" % (k, got[k], want[k]) for k in diff[:4]]
    raise AssertionError("%s wires differ: %d keys only in first %s, %d only in second %s, %d with different lengths %s"
                         % (what, len(only_a), only_a[:5], len(only_b), only_b[:5], len(diff), detail))
