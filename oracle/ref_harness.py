"""TEST INFRASTRUCTURE - live-reference harness (not part of the product path).

Imports the reference's own ``mdhbond`` package, unmodified, from ``$CGRAPHS_REF`` (default
``/root/reference/cgraphs``) so that its ``HbondAnalysis`` can be executed in this container and used
to (a) validate the CPU restatement in ``oracle/hbond_port.py`` and (b) generate the golden vectors
under ``tests/golden/`` (``tests/golden/make_golden.py``).

The reference hard-imports ``MDAnalysis`` and ``matplotlib`` (``cgraphs/mdhbond/basic.py:24-30``,
``hbond.py:25-35``, ``network.py:24-34``); neither is installable offline.  They are
replaced in ``sys.modules`` by thin stubs: ``MDAnalysis.Universe(structure)`` returns the
``cgraphs_b200.universe.Universe`` it is given, ``MDAnalysis.core.groups.AtomGroup`` is the array
backed AtomGroup of that module, and the matplotlib names only need to exist.

``/root/reference`` does not exist on the GPU box; there ``CGRAPHS_REF`` points at ``oracle/_ref/cgraphs``, the
reference's own modules byte-compiled by ``oracle/make_ref.py`` (compiler output, git-ignored, travels with the
snapshot). Only ``tests/`` (CPU side), ``tests/golden/make_golden.py`` and ``bench.py --impl reference`` import this
module.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_DEFAULT = "/root/reference/cgraphs"


def reference_available(path=None):
    path = path or os.environ.get("CGRAPHS_REF", REF_DEFAULT)
    return os.path.isfile(os.path.join(path, "mdhbond", "hbond.py")) or os.path.isfile(os.path.join(path, "mdhbond", "hbond.pyc.bin"))


def _install_stubs():
    from cgraphs_b200 import universe as uni

    if "MDAnalysis" not in sys.modules or getattr(sys.modules["MDAnalysis"], "_hbgpu_stub", False):
        mda = types.ModuleType("MDAnalysis")
        mda._hbgpu_stub = True

        def Universe(structure, trajectories=None, *a, **k):
            if isinstance(structure, uni.Universe):
                return structure
            raise RuntimeError("stub MDAnalysis: pass a cgraphs_b200.universe.Universe as `structure`")

        mda.Universe = Universe
        core = types.ModuleType("MDAnalysis.core")
        groups = types.ModuleType("MDAnalysis.core.groups")
        groups.AtomGroup = uni.AtomGroup
        ag_mod = types.ModuleType("MDAnalysis.core.AtomGroup")
        ag_mod.AtomGroup = uni.AtomGroup
        core.groups = groups
        core.AtomGroup = ag_mod
        mda.core = core
        mda.AtomGroup = uni.AtomGroup
        sys.modules["MDAnalysis"] = mda
        sys.modules["MDAnalysis.core"] = core
        sys.modules["MDAnalysis.core.groups"] = groups
        sys.modules["MDAnalysis.core.AtomGroup"] = ag_mod
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl._hbgpu_stub = True
        plt = types.ModuleType("matplotlib.pyplot")
        ticker = types.ModuleType("matplotlib.ticker")
        ticker.MaxNLocator = type("MaxNLocator", (), {"__init__": lambda self, *a, **k: None})
        mpl.pyplot = plt
        mpl.ticker = ticker
        mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.ticker"] = ticker


_REF = None


def load_reference(path=None):
    """Return the reference's ``mdhbond`` package (imported once)."""
    global _REF
    if _REF is not None:
        return _REF
    path = path or os.environ.get("CGRAPHS_REF", REF_DEFAULT)
    if not reference_available(path):
        raise ImportError("reference mdhbond not found under %s" % path)
    _install_stubs()
    # import `mdhbond` as a top-level package, skipping cgraphs/__init__.py (tkinter, Bio, sklearn GUI imports)
    init = os.path.join(path, "mdhbond", "__init__.py")
    if not os.path.isfile(init):          # oracle/_ref: the reference byte-compiled by oracle/make_ref.py (*.pyc.bin)
        import shutil
        import tempfile
        tmp = tempfile.mkdtemp(prefix="cgraphs_ref_")
        for root, _, files in os.walk(path):
            for fn in files:
                if fn.endswith(".pyc.bin"):
                    dst = os.path.join(tmp, os.path.relpath(root, path), fn[:-4])
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    shutil.copyfile(os.path.join(root, fn), dst)
        path = tmp
        init = os.path.join(path, "mdhbond", "__init__.pyc")
    spec = importlib.util.spec_from_file_location(
        "_cgraphs_ref_mdhbond", init, submodule_search_locations=[os.path.join(path, "mdhbond")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_cgraphs_ref_mdhbond"] = mod
    spec.loader.exec_module(mod)
    _REF = mod
    return mod


def run_reference(universe, selection, method="in_selection", around=3.5, exclude_backbone_backbone=True,
                  **ctor_kwargs):
    """Run the reference HbondAnalysis on a stub Universe and return the analysis object."""
    ref = load_reference()
    hba = ref.HbondAnalysis(selection, universe, **ctor_kwargs)
    if method == "in_selection":
        hba.set_hbonds_in_selection(exclude_backbone_backbone=exclude_backbone_backbone)
    elif method == "water_around":
        hba.set_hbonds_in_selection_and_water_around(around, exclude_backbone_backbone=exclude_backbone_backbone)
    else:
        raise ValueError(method)
    return hba


def results_bytes(results):
    """Canonical comparable form of a result dict (SURVEY 8c parity definition)."""
    return {k: bytes(memoryview(v.astype(bool))) for k, v in results.items()}


_REF_WORKFLOW = None


def load_reference_workflow(path=None):
    """The reference's ``conservedgraph`` module (``ConservedGraph.get_conserved_graph``, conservedgraph.py:56-170),
    imported under a throw-away package so that ``cgraphs/__init__.py`` (GUI imports) is never executed. Needs the
    sources (this container): ``Bio``, ``MDAnalysis.analysis`` and ``matplotlib.cm`` are stubbed - the function under
    test uses none of them."""
    global _REF_WORKFLOW
    if _REF_WORKFLOW is not None:
        return _REF_WORKFLOW
    path = path or os.environ.get("CGRAPHS_REF", REF_DEFAULT)
    if not os.path.isfile(os.path.join(path, "conservedgraph.py")):
        raise ImportError("reference conservedgraph.py not found under %s" % path)
    _install_stubs()
    mda = sys.modules["MDAnalysis"]
    if "MDAnalysis.analysis" not in sys.modules:
        ana = types.ModuleType("MDAnalysis.analysis")
        dist = types.ModuleType("MDAnalysis.analysis.distances")
        ana.distances = dist
        mda.analysis = ana
        sys.modules["MDAnalysis.analysis"] = ana
        sys.modules["MDAnalysis.analysis.distances"] = dist
    if "Bio" not in sys.modules:
        bio = types.ModuleType("Bio")
        svd = types.ModuleType("Bio.SVDSuperimposer")
        svd.SVDSuperimposer = type("SVDSuperimposer", (), {})
        align = types.ModuleType("Bio.Align")
        bio.SVDSuperimposer = svd
        bio.Align = align
        sys.modules["Bio"] = bio
        sys.modules["Bio.SVDSuperimposer"] = svd
        sys.modules["Bio.Align"] = align
    mpl = sys.modules["matplotlib"]
    if getattr(mpl, "_hbgpu_stub", False) and "matplotlib.cm" not in sys.modules:
        cm = types.ModuleType("matplotlib.cm")
        mpl.cm = cm
        sys.modules["matplotlib.cm"] = cm
        for name in ("lines", "patches", "colors"):
            m = types.ModuleType("matplotlib." + name)
            setattr(mpl, name, m)
            sys.modules["matplotlib." + name] = m
    pkg = types.ModuleType("_cgraphs_ref_pkg")
    pkg.__path__ = [path]
    sys.modules["_cgraphs_ref_pkg"] = pkg
    mod = importlib.import_module("_cgraphs_ref_pkg.conservedgraph")
    _REF_WORKFLOW = mod
    return mod


def reference_conserved_graph(graph_coord_objects, reference_coordinates, conservation_threshold=0.9, eps=1.5):
    """Run the reference's ConservedGraph.get_conserved_graph (graph_type "hbond") on prepared per-structure objects
    ({name: {"graph": networkx graph, "structure": Universe}}) without its file-based constructor."""
    import logging
    mod = load_reference_workflow()
    obj = mod.ConservedGraph.__new__(mod.ConservedGraph)
    obj.logger = logging.getLogger("cgraphs_ref_stub")
    obj.graph_type = "hbond"
    obj.graph_coord_objects = graph_coord_objects
    obj.reference_coordinates = reference_coordinates
    obj.get_conserved_graph(conservation_threshold=conservation_threshold, eps=eps)
    return obj.conserved_nodes, obj.conserved_edges
