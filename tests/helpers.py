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
