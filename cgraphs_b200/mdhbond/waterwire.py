"""``WireAnalysis`` - water wires on the GPU (SURVEY 8f "next" #1; reference ``cgraphs/mdhbond/waterwire.py``).

cgraphs' trajectory workflow (``cgraphs/proteingraphanalyser.py:472-485``) builds one ``WireAnalysis`` per
trajectory, calls ``set_water_wires(max_water=...)``, ``compute_average_water_per_wire()`` and reads
``filtered_graph``. This class mirrors that surface: same constructor keywords as the reference
(``waterwire.py:42-45``), ``set_water_wires`` (``waterwire.py:191-317``) and the results it publishes -
``initial_results`` / ``filtered_results`` (bool per frame), ``wire_lengths`` (float per frame: inf = no wire,
0 = direct H-bond, k = waters in the shortest wire) and the graphs.

The neighbour stage (``waterwire.py:200-248``) is the library's water-around search with the radius
``(max_water + 1) * distance / 2``; the per-frame breadth-first search (``waterwire.py:250-309``) runs in
``k_wire_bfs`` (``csrc/hb_wires.cuh``).

``water_in_convex_hull`` (what cgraphs itself always passes, ``proteingraphanalyser.py:484``): the local waters of every
frame come from the GPU (the water-presence search), the Delaunay test of the reference (``waterwire.py:218-221``,
scipy / Qhull - a library call the reference makes too, not restated in CUDA) runs on the host frame by frame, and the
per-frame lists "position of the j-th in-hull water, named like the j-th local water" - the reference does not filter
its index list along with the coordinates, ``waterwire.py:221-228`` - go back to the device, where ``k_virtual_pairs``
searches them and the same breadth-first search follows. Slow (a Delaunay triangulation per frame, as in the
reference) but drop-in: the results are the reference's, labelling included.

Not covered (out of scope): the explicit wire paths (``hashs`` / ``hash_table``: Python ``hash(str(path))``
values, salted per process) and the plotting methods.
"""
from __future__ import annotations

import numpy as _np

from .hbond import HbondAnalysis


def split_wire_rows(words, n_frames, frame0=0):
    """Rows of a wire run (include/hbgpu.h, hb_set_wires) -> (uint8 [n, n_frames] per-frame bytes,
    uint64 [n] `(first direct frame << 32) | (acceptor entry << 1) | donor_is_second_node`, all ones = never)."""
    n = len(words)
    if n == 0:
        return _np.zeros((0, n_frames), _np.uint8), _np.zeros(0, _np.uint64)
    nb4 = words.shape[1] - 2
    cols = _np.ascontiguousarray(words[:, :nb4].view(_np.uint8)[:, :n_frames])
    tail = _np.ascontiguousarray(words[:, nb4:nb4 + 2]).view(_np.uint64)[:, 0]
    packed = ~tail                                   # the device keeps the complement (atomicMax on zeroed rows)
    has = tail != 0
    packed = _np.where(has, packed + (_np.uint64(frame0) << _np.uint64(32)), _np.uint64(0xFFFFFFFFFFFFFFFF))
    return cols, packed


def decode_wire_rows(keys, cols, first_direct, node_entry, ids, n_frames):
    """(node_lo << 32 | node_hi) keys + per-frame bytes -> the reference's `wire_lengths` dict
    (waterwire.py:286-293, 302-309): float64 per frame, inf = no wire, 0 = direct H-bond, k = waters in the wire.
    Key orientation = first discovery (waterwire.py:277-279, 299-301): a wire names the residue with the smaller
    node first; a direct bond names the donor's residue first; within a frame wires are met first."""
    results = {}
    if len(keys) == 0:
        return results
    wire_k = cols & 0x7f
    direct = (cols & 0x80) != 0
    lengths = _np.full(cols.shape, _np.inf)
    has = wire_k > 0
    lengths[has] = wire_k[has]
    lengths[direct] = 0
    has_wire = has.any(axis=1)
    first_wire = _np.where(has_wire, has.argmax(axis=1), n_frames)
    has_direct = first_direct != _np.uint64(0xFFFFFFFFFFFFFFFF)
    fd_frame = _np.where(has_direct, (first_direct >> _np.uint64(32)).astype(_np.int64), n_frames)
    donor_hi = (first_direct & _np.uint64(1)).astype(bool) & has_direct
    by_wire = has_wire & (first_wire <= fd_frame)
    lo = node_entry[(keys >> _np.uint64(32)).astype(_np.int64)]
    hi = node_entry[(keys & _np.uint64(0xFFFFFFFF)).astype(_np.int64)]
    swap = ~by_wire & donor_hi
    a = _np.where(swap, hi, lo)
    b = _np.where(swap, lo, hi)
    order = _np.lexsort((_np.arange(len(keys)), _np.minimum(first_wire, fd_frame)))
    for i in order:
        results[ids[a[i]] + ':' + ids[b[i]]] = lengths[i]
    return results


class WireAnalysis(HbondAnalysis):

    def __init__(self, selection=None, structure=None, trajectories=None, distance=3.5, cut_angle=60.,
                 start=None, stop=None, step=1, additional_donors=[], additional_acceptors=[], exclude_donors=[],
                 exclude_acceptors=[], ions=[], check_angle=True, residuewise=True,
                 add_donors_without_hydrogen=False, restore_filename=None, **kw):
        self.hashs = {}
        self.hash_table = {}
        self.wire_lengths = {}
        super().__init__(selection=selection, structure=structure, trajectories=trajectories, distance=distance,
                         cut_angle=cut_angle, start=start, stop=stop, step=step, additional_donors=additional_donors,
                         additional_acceptors=additional_acceptors, exclude_donors=exclude_donors,
                         exclude_acceptors=exclude_acceptors, ions=ions, check_angle=check_angle,
                         residuewise=residuewise, add_donors_without_hydrogen=add_donors_without_hydrogen,
                         restore_filename=restore_filename, **kw)
        if restore_filename is not None:
            return
        top = self._topology
        # waterwire.py:60-65: `da_trans` - all entries with one residue-level id are the first of them; wires are booked
        # on that first entry whatever `residuewise` says, direct bonds on the entries themselves (waterwire.py:296-301).
        # Nodes: residuewise - one per residue; else one per atom-level id (the donor and the acceptor entry of one atom
        # share their id string, so they share a node).
        res_ids = ['-'.join(s.split('-')[:3]) for s in top.all_ids_atomwise[:top.first_water]]
        first = {}
        da_trans = _np.empty(top.first_water, dtype=_np.int64)
        for e, rid in enumerate(res_ids):
            da_trans[e] = first.setdefault(rid, e)
        node_ids = res_ids if self.residuewise else list(top.all_ids_atomwise[:top.first_water])
        node_of_id = {}
        node_first_entry = []
        own_node = _np.empty(top.first_water, dtype=_np.int32)
        for e, nid in enumerate(node_ids):
            n = node_of_id.get(nid)
            if n is None:
                n = node_of_id[nid] = len(node_first_entry)
                node_first_entry.append(e)
            own_node[e] = n
        self._n_wire_nodes = len(node_first_entry)
        ent_node = _np.empty(top.first_water + top.nW, dtype=_np.int32)
        ent_node[:top.first_water] = own_node[da_trans]
        ent_node[top.first_water:] = self._n_wire_nodes + _np.arange(top.nW, dtype=_np.int32)
        self._wire_ent_node = ent_node
        self._wire_direct_node = None
        if not self.residuewise:
            direct = ent_node.copy()
            direct[:top.first_water] = own_node
            self._wire_direct_node = direct
        self._wire_node_entry = _np.asarray(node_first_entry, dtype=_np.int64)
        self.da_trans = da_trans

    def _hull_water_lists(self, max_water, lo, hi):
        """Per frame of [lo, hi): (offsets, water points whose positions are searched, water points they are named as) -
        waterwire.py:212-228 with `water_in_convex_hull` set, see the module docstring."""
        from scipy.spatial import Delaunay
        from .hbond import _frame_blocks
        top = self._topology
        if (top.sel_wat >= 0).any():
            raise NotImplementedError('water_in_convex_hull with waters of the water group inside the selection')
        radius = float(max_water + 1) * self.distance / 2.
        keys, words = self._run("presence", radius, False)              # GPU: local waters of every frame
        n_local = hi - lo
        present = _np.zeros((top.nW, n_local), dtype=bool)
        if len(keys):
            bits = _np.unpackbits(words.view(_np.uint8), axis=1, bitorder="little")[:, :n_local].astype(bool)
            present[(keys & _np.uint64(0xFFFFFFFF)).astype(_np.int64) - top.first_water] = bits
        sel_atoms = top.entry_atom[:top.first_water]                     # donors ++ acceptors, like _da_selection
        water_atoms = top.water_atoms
        off, true, label = [0], [], []
        block = max(1, int((256 << 20) // max(1, self._n_atoms * 12)))
        col = 0
        for xyz in _frame_blocks(self._universe, self._frame_indices[lo:hi], block):
            for pos in xyz:
                local = _np.nonzero(present[:, col])[0]
                if len(local):
                    hull = Delaunay(pos[sel_atoms])                      # waterwire.py:219
                    inside = hull.find_simplex(pos[water_atoms[local]]) != -1
                    t = local[inside]
                    true.append(t)
                    label.append(local[:len(t)])                         # (the unfiltered list keeps naming them)
                    off.append(off[-1] + len(t))
                else:
                    off.append(off[-1])
                col += 1
        cat = lambda parts: _np.concatenate(parts).astype(_np.int32) if parts else _np.zeros(0, _np.int32)
        return _np.asarray(off, dtype=_np.int64), cat(true), cat(label)

    def set_water_wires(self, max_water=5, allow_direct_bonds=True, water_in_convex_hull=False):
        """waterwire.py:191-317."""
        max_water = int(max_water)
        self._allow_direct_bonds = allow_direct_bonds
        around = float(max_water + 1) * self.distance / 2.
        wires = (self._wire_ent_node, self._n_wire_nodes, max_water, bool(allow_direct_bonds))
        if water_in_convex_hull:
            if self._pbc:
                raise NotImplementedError('water_in_convex_hull has no minimum-image form')
            lo, hi = 0, self.nb_frames
            if self._shard is not None:
                from .sharding import shard_range
                lo, hi = shard_range(self.nb_frames, *self._shard)
            lists = self._hull_water_lists(max_water, lo, hi)
            keys, words = self._run("in_selection", 0.0, False, wires=wires + (lists,))
        else:
            keys, words = self._run("water_around", around, False, wires=wires)
        lo, hi = 0, self.nb_frames
        if self._shard is not None:
            from .sharding import shard_range
            lo, hi = shard_range(self.nb_frames, *self._shard)
        cols, tail = split_wire_rows(words, hi - lo, lo)
        if self._shard is not None and self._shard[1] > 1:
            from .sharding import merge_wire_shards
            keys, cols, tail = merge_wire_shards(keys, cols, tail, lo, hi, self.nb_frames)
        results = decode_wire_rows(keys, cols, tail, self._wire_node_entry, self._all_ids, self.nb_frames)
        self._set_results({key: results[key] != _np.inf for key in results})
        self.wire_lengths = results
        self.hashs = {}
        self.hash_table = {}
        self.occupancy_dict = {}
        self.first_frame_dict = {}

    def compute_average_water_per_wire(self, use_filtered=True):
        """waterwire.py:359-362."""
        results = self.filtered_results if use_filtered else self.initial_results
        return {key: self.wire_lengths[key][self.wire_lengths[key] < _np.inf].mean() for key in results}
