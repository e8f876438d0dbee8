"""The CPU restatement (oracle/hbond_port.py) reproduces the golden vectors that the live reference
produced (tests/golden/make_golden.py) - with the cKDTree back-end, the brute-force DIST predicate,
and the cosine-threshold form of the angle test the kernels use."""
import pytest

from oracle import hbond_port
from tests.helpers import assert_same_results, golden_cases, graph_edges, lists_to_results, load_golden

_CACHE = {}


def _fixture(name):
    if name not in _CACHE:
        _CACHE[name] = load_golden(name)
    return _CACHE[name]


@pytest.mark.parametrize("name,idx", golden_cases())
def test_port_matches_golden(name, idx):
    u, runs = _fixture(name)
    run = runs[idx]
    expect = lists_to_results(run["results"], run["nb_frames"])
    big = run["selection"] != "protein" and name == "peptide_traj"
    for backend, cthr in (("kdtree", None), ("brute", run["cos_threshold"])):
        if backend == "brute" and big and idx % 4:
            continue   # keep the CPU suite short; the brute predicate is covered by the other cases
        res = hbond_port.run(u, run["selection"], run["method"], around=run["around"],
                             exclude_backbone_backbone=run["excl_bb"], backend=backend, cos_threshold=cthr,
                             **run["kwargs"])
        assert res.nb_frames == run["nb_frames"]
        assert_same_results(res.initial_results, expect, "%s[%d] %s" % (name, idx, backend))
        assert graph_edges(res.filtered_graph) == {frozenset(e) for e in run["graph_edges"]}
