"""Batched static structures: many independent structures searched in one launch (BASELINE config C5).

cgraphs' conserved-graph workflow builds one ``HbondAnalysis`` per PDB structure and runs it on a
single frame (reference ``cgraphs/proteingraphanalyser.py:370-450``).  ``HbondBatch`` does the same
work for a list of structures at once: the topologies are concatenated into one ragged ``hb_topology``
(one system per structure), the coordinates into one ragged buffer, and "frame" f of the occupancy
matrix is structure f.  Bond keys share one id space over the batch, so the matrix is directly the
bond x structure table the conserved-graph reduction needs (``conservedgraph.py:114-170``).
"""
from __future__ import annotations

import numpy as _np

from . import _native
from .calib import cos_threshold as _cos_threshold
from .hbond import HbondAnalysis, dict2graph
from .topology import pack_systems


class StructureResult(object):
    """What a per-structure ``HbondAnalysis`` exposes after a ``set_hbonds_*`` call."""

    def __init__(self, results, residuewise, graph=None):
        self.initial_results = results
        self.filtered_results = results
        self.nb_frames = 1
        self.residuewise = residuewise
        self.initial_graph = dict2graph(results, residuewise) if graph is None else graph
        self.filtered_graph = self.initial_graph


WATER_TYPES = ("HOH", "TIP3", "TIP4")        # cgraphs/helperfunctions.py:47


class HbondBatch(object):
    def __init__(self, structures, selection, device=None, native_options=None, **ctor_kwargs):
        """structures: Universe-like objects (single frame each); other kwargs as HbondAnalysis."""
        if not structures:
            raise AssertionError('No structure file path.')
        self._device = device
        self._native_options = dict(native_options or {})
        self.analyses = [HbondAnalysis(selection, s, device=device, **ctor_kwargs) for s in structures]
        a0 = self.analyses[0]
        self.check_angle, self.residuewise = a0.check_angle, a0.residuewise
        self.distance, self.cut_angle = a0.distance, a0.cut_angle
        self._ctx = None
        self.results = None
        self.bond_keys = None
        self.occupancy_words = None
        self.last_run_stats = None

    def set_hbonds_in_selection(self, exclude_backbone_backbone=True):
        for a in self.analyses:
            if a._nb_acceptors == 0: raise AssertionError('No acceptors in the selection')
            if a._nb_donors == 0: raise AssertionError('No donors in the selection')
        return self._run("in_selection", 0.0, exclude_backbone_backbone)

    def set_hbonds_in_selection_and_water_around(self, around_radius, exclude_backbone_backbone=True):
        return self._run("water_around", float(around_radius), exclude_backbone_backbone)

    def _run(self, method, around, excl_bb):
        tops = [a._topology for a in self.analyses]
        for a in self.analyses:
            a._check_hydrogen_tables(method)
        n_atoms = [a._n_atoms for a in self.analyses]
        cache = self.__dict__.setdefault("_packed", {})
        if method not in cache:
            cache[method] = pack_systems(tops, [t.key_ids(method) for t in tops], n_atoms)
        tables, group_names = cache[method]
        if self._ctx is None:
            self._ctx = _native.Context(self._device if self._device is not None else 0)
            for name, value in self._native_options.items():
                self._ctx.set_option(name, value)
        ctx = self._ctx
        if getattr(ctx, "_uploaded", None) is not tables:
            ctx.set_topology(tables)
            ctx._uploaded = tables
        cthr = _cos_threshold(float(self.cut_angle)) if self.check_angle else -1.0
        ctx.set_params(self.distance, around, cthr, self.check_angle, excl_bb,
                       _native.METHOD_WATER_AROUND if method == "water_around" else _native.METHOD_IN_SELECTION,
                       structure_batch=True)
        n = len(self.analyses)
        xyz = _np.concatenate([a._first_frame_positions() for a in self.analyses]).astype(_np.float32)
        ctx.begin(n)
        ctx.submit_raw(xyz.ctypes.data, 0, n)
        keys, words = ctx.results()
        self.last_run_stats = ctx.timers()
        self.bond_keys = [group_names[int(k >> _np.uint64(32))] + ':' + group_names[int(k & _np.uint64(0xFFFFFFFF))]
                          for k in keys]
        self.occupancy_words = words
        bits = _np.unpackbits(words.view(_np.uint8), axis=1, bitorder="little")[:, :n].astype(bool) \
            if len(keys) else _np.zeros((0, n), bool)
        self.occupancy = bits
        self.results = []
        for i, a in enumerate(self.analyses):
            rows = _np.nonzero(bits[:, i])[0]
            res = {self.bond_keys[r]: _np.ones(1, dtype=bool) for r in rows}
            graph = dict2graph(res, self.residuewise)      # one graph per structure, shared by the views below
            a.initial_results = a.filtered_results = res   # (what _set_results does, basic.py:94-98)
            a.initial_graph = a.filtered_graph = graph
            self.results.append(StructureResult(res, self.residuewise, graph))
        return self.results

    # ---------------------------------------------------------------------------------------------
    # conserved network over the batch (SURVEY 8f "next" #3): the counting rule of
    # ConservedGraph.get_conserved_graph for graph_type "hbond" (cgraphs/conservedgraph.py:114-170, without the
    # water-cluster relabelling of :139-149, which needs the workflow's superimposed reference coordinates),
    # evaluated on the bond x structure occupancy matrix instead of one networkx graph per structure
    # ---------------------------------------------------------------------------------------------
    def get_conserved_graph(self, conservation_threshold=0.9):
        """Returns (conserved_nodes, conserved_edges): nodes / edges present in at least
        round(n_structures * conservation_threshold) of the per-structure graphs. Edges are (node, node) tuples with
        the smaller resid first (the orientation dict2graph gives them, helpfunctions.py:120-134); water nodes are
        kept only when they take part in a conserved edge (conservedgraph.py:160-168)."""
        if self.results is None:
            raise AssertionError('run set_hbonds_in_selection / set_hbonds_in_selection_and_water_around first')
        n = len(self.analyses)
        th = _np.round(n * conservation_threshold)
        cut = 3 if self.residuewise else 4
        edge_of_row = _np.full(len(self.bond_keys), -1, dtype=_np.int64)
        edges, index = [], {}
        for r, key in enumerate(self.bond_keys):                 # graph edge of every bond key (once per key)
            a, b = key.split(':')
            a = '-'.join(a.split('-')[:cut])
            b = '-'.join(b.split('-')[:cut])
            if int(a.split('-')[2]) > int(b.split('-')[2]):
                a, b = b, a
            if a == b:
                continue
            e = frozenset((a, b))
            if e not in index:
                index[e] = len(edges)
                edges.append((a, b))
            edge_of_row[r] = index[e]
        occ = self.occupancy                                     # bool [n_bonds, n_structures]
        keep = edge_of_row >= 0
        edge_present = _np.zeros((len(edges), n), dtype=bool)
        _np.logical_or.at(edge_present, edge_of_row[keep], occ[keep])
        edge_counts = edge_present.sum(axis=1)
        node_index, node_names = {}, []
        ends = _np.empty((len(edges), 2), dtype=_np.int64)
        for i, (a, b) in enumerate(edges):
            for j, name in enumerate((a, b)):
                if name not in node_index:
                    node_index[name] = len(node_names)
                    node_names.append(name)
                ends[i, j] = node_index[name]
        node_present = _np.zeros((len(node_names), n), dtype=bool)
        if len(edges):
            _np.logical_or.at(node_present, ends[:, 0], edge_present)
            _np.logical_or.at(node_present, ends[:, 1], edge_present)
        node_counts = node_present.sum(axis=1)
        conserved_edges = [edges[i] for i in _np.nonzero(edge_counts >= th)[0]]
        in_edge = {x for e in conserved_edges for x in e}
        conserved_nodes = [node_names[i] for i in _np.nonzero(node_counts >= th)[0]
                           if node_names[i].split('-')[1] not in WATER_TYPES or node_names[i] in in_edge]
        self.conserved_edges = sorted(conserved_edges)
        self.conserved_nodes = sorted(conserved_nodes)
        return self.conserved_nodes, self.conserved_edges
