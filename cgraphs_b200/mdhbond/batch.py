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
    # conserved network over the batch (SURVEY 8f "next" #3): ConservedGraph.get_conserved_graph for graph_type "hbond"
    # (cgraphs/conservedgraph.py:114-170) evaluated on the bond x structure occupancy matrix instead of one networkx
    # graph per structure, including the relabelling of waters that sit on a conserved-water cluster centre
    # (conservedgraph.py:118-149).
    # ---------------------------------------------------------------------------------------------
    def _graph_edges_of_rows(self):
        """Graph edge (a, b) of every bond key like dict2graph forms it (helpfunctions.py:120-134); -1: self loop."""
        cut = 3 if self.residuewise else 4
        edge_of_row = _np.full(len(self.bond_keys), -1, dtype=_np.int64)
        edges, index = [], {}
        for r, key in enumerate(self.bond_keys):
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
        return edges, edge_of_row

    def _water_positions(self, i):
        """conservedgraph.py:118-124: resid -> position of the waters of structure i (a later water with the same resid
        replaces an earlier one, as in the reference's dict)."""
        a = self.analyses[i]
        top = a._topology
        pos = a._first_frame_positions()
        cols_resid = _np.asarray(a._universe.atoms.resids)
        out = {}
        for atom in top.water_atoms:
            out[int(cols_resid[atom])] = pos[atom]
        return out

    def get_conserved_graph(self, conservation_threshold=0.9, eps=1.5, reference_coordinates=None):
        """Returns (conserved_nodes, conserved_edges): nodes / edges present in at least
        round(n_structures * conservation_threshold) of the per-structure graphs (conservedgraph.py:151-169).

        reference_coordinates: the workflow's ``{'X-w-<k>': (x, y, z), ...}`` table of conserved-water cluster centres
        (conservedgraph.py:41-45); a water end of an edge that lies within ``eps`` of a centre is renamed to that
        centre's node before the edges are counted (conservedgraph.py:131-149; the last matching centre wins, nodes are
        counted under their own names). Edges are (node, node) tuples, unordered: the reference orients a conserved edge
        like its first occurrence in networkx's iteration order, which is not a property of the H-bond data.
        Water nodes are kept only when they take part in a conserved edge (conservedgraph.py:160-168).

        Without water clusters the edge counts are the per-row popcounts of the packed occupancy matrix that is still
        on the device (hb_get_counts); rows whose edge may be renamed are counted structure by structure."""
        if self.results is None:
            raise AssertionError('run set_hbonds_in_selection / set_hbonds_in_selection_and_water_around first')
        n = len(self.analyses)
        th = _np.round(n * conservation_threshold)
        edges, edge_of_row = self._graph_edges_of_rows()
        occ = self.occupancy                                     # bool [n_bonds, n_structures]
        keep = edge_of_row >= 0
        is_water = lambda name: name.split('-')[1] in WATER_TYPES
        centres = [(k, _np.asarray(c, dtype=_np.float64)) for k, c in (reference_coordinates or {}).items() if k.startswith("X-w")]
        rows_per_edge = _np.bincount(edge_of_row[keep], minlength=len(edges))
        counts = {}                                              # unordered edge -> instances over the structures
        # edges that cannot be renamed and come from exactly one row: device popcount of that row
        simple = _np.array([rows_per_edge[i] == 1 and not (centres and (is_water(a) or is_water(b)))
                            for i, (a, b) in enumerate(edges)], dtype=bool) if edges else _np.zeros(0, bool)
        row_counts = None
        if simple.any() and self._ctx is not None and getattr(self._ctx, "_run_id", None) is not None:
            try:
                row_counts = self._ctx.counts(len(self.bond_keys))
            except Exception:
                row_counts = None
        if row_counts is None:
            row_counts = occ.sum(axis=1)
        for r in _np.nonzero(keep)[0]:
            e = edge_of_row[r]
            if simple[e]:
                counts[frozenset(edges[e])] = counts.get(frozenset(edges[e]), 0) + int(row_counts[r])
        # the rest: per structure, the distinct graph edges present, water ends renamed by that structure's coordinates
        hard = _np.nonzero(~simple)[0] if len(edges) else []
        if len(hard):
            hard_rows = _np.nonzero(keep & ~simple[_np.where(keep, edge_of_row, 0)])[0]
            present = _np.zeros((len(edges), n), dtype=bool)
            _np.logical_or.at(present, edge_of_row[hard_rows], occ[hard_rows])
            eps2 = float(eps) ** 2
            for i in range(n):
                idx = _np.nonzero(present[:, i])[0]
                if not len(idx):
                    continue
                waters = self._water_positions(i) if centres else {}
                for e in idx:
                    ends = list(edges[e])
                    for j in (0, 1):
                        if centres and is_water(ends[j]):
                            w = waters.get(int(ends[j].split('-')[2]))
                            if w is not None:
                                for key, cc in centres:
                                    d = _np.asarray(w, dtype=_np.float64) - cc
                                    if d[0] ** 2 + d[1] ** 2 + d[2] ** 2 <= eps2:
                                        ends[j] = "X-w-" + key.split("-")[-1]
                    k = frozenset(ends) if ends[0] != ends[1] else (ends[0],)
                    counts[k] = counts.get(k, 0) + 1
        # nodes under their own names: a node is in a structure's graph iff one of its edges is
        edge_present = _np.zeros((len(edges), n), dtype=bool)
        if keep.any():
            _np.logical_or.at(edge_present, edge_of_row[keep], occ[keep])
        node_index, node_names = {}, []
        ends_ix = _np.empty((len(edges), 2), dtype=_np.int64)
        for i, (a, b) in enumerate(edges):
            for j, name in enumerate((a, b)):
                if name not in node_index:
                    node_index[name] = len(node_names)
                    node_names.append(name)
                ends_ix[i, j] = node_index[name]
        node_present = _np.zeros((len(node_names), n), dtype=bool)
        if len(edges):
            _np.logical_or.at(node_present, ends_ix[:, 0], edge_present)
            _np.logical_or.at(node_present, ends_ix[:, 1], edge_present)
        node_counts = node_present.sum(axis=1)
        conserved_edges = []
        for k, c in counts.items():
            if c >= th:
                t = tuple(sorted(k)) if isinstance(k, frozenset) else (k[0], k[0])
                conserved_edges.append(t)
        in_edge = {x for e in conserved_edges for x in e}
        conserved_nodes = [node_names[i] for i in _np.nonzero(node_counts >= th)[0]
                           if not is_water(node_names[i]) or node_names[i] in in_edge]
        self.conserved_edges = sorted(conserved_edges)
        self.conserved_nodes = sorted(conserved_nodes)
        self.conserved_edge_counts = {tuple(sorted(k)) if isinstance(k, frozenset) else (k[0], k[0]): c for k, c in counts.items()}
        return self.conserved_nodes, self.conserved_edges
