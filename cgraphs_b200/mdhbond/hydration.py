"""``HydrationAnalysis.set_presence_around`` on the GPU (SURVEY 8f "next" #4; reference
``cgraphs/mdhbond/hydration.py:71-107``).

Per frame: the waters of the water group (``basic.py:62``) within ``around_distance`` of ANY atom of the selection
(all selected atoms, not only donors / acceptors). That is the local-water test of the H-bond path
(``hbond.py:249-252``) with the whole selection as the point set, so the library runs the first half of its
water-around pipeline (``HB_METHOD_WATER_PRESENCE``: selection cell grid, ``k_local_water``) and sets one bit per
present water and frame (``k_presence_emit``). Results: ``initial_results`` / ``filtered_results`` =
``{water id: bool[nb_frames]}`` like the reference's, plus ``filter_occupancy`` (``hydration.py:109-115``).

``per_residue=True`` (``hydration.py:88-94``: one key ``'residue id:water id'`` per selected atom - water pair within
the distance) runs the distance-only pair search of the water-around method with ``distance = around_distance``: every
selected atom is an acceptor-like entry whose key group is its residue id, every water a water entry; the water-water
pairs the search also finds are dropped on the host. Selections that contain waters of the water group are not
supported in this mode (the reference's key orientation then depends on discovery order).

Not covered: the convex-hull variant ``set_presence_in_hull`` (scipy Delaunay) and the diffusion / residence-time
statistics further down in the reference class.
"""
from __future__ import annotations

import numpy as _np

from . import _native
from .hbond import WATER_DEFINITION, _as_universe, _frame_blocks
from .topology import AtomColumns, _ids


class HydrationAnalysis(object):

    def __init__(self, selection=None, structure=None, trajectories=None, start=None, stop=None, step=1,
                 restore_filename=None, device=None, native_options=None):
        if restore_filename is not None:
            raise NotImplementedError('restore_filename is not supported by the GPU HydrationAnalysis')
        # basic.py:41-51
        if selection is None: raise AssertionError('No selection string.')
        if structure is None: raise AssertionError('No structure file path.')
        self._selection = selection
        self._universe = _as_universe(structure, trajectories)
        self._trajectory_slice = slice(start if isinstance(start, int) else None,
                                       stop if isinstance(stop, int) else None, step)
        self._mda_selection = self._universe.select_atoms(selection)
        if not len(self._mda_selection): raise AssertionError('No atoms match the selection')
        self._water = self._universe.select_atoms(WATER_DEFINITION)
        cols = AtomColumns(self._universe)
        self._n_atoms = cols.n_atoms
        self._water_ids = list(_ids(cols, _np.asarray(self._water.indices), False))
        self._water_ids_atomwise = list(_ids(cols, _np.asarray(self._water.indices), True))
        self.initial_results = {}
        self.filtered_results = {}
        n_traj = len(self._universe.trajectory)
        self._frame_indices = range(*self._trajectory_slice.indices(n_traj))
        self.nb_frames = len(self._frame_indices)
        self.applied_filters = {'resnames': None, 'segnames': None, 'shortest_paths': None, 'single_path': None,
                                'occupancy': None, 'connected_component': None, 'shells': None,
                                'avg_least_bonds': None, 'backbone': None}
        self._device = device
        self._native_options = dict(native_options or {})
        self._ctx = None
        self._cols = cols
        self.last_run_stats = None

    def _tables(self):
        """hb_topology arrays: every selected atom is a selection point without entries, every water one entry."""
        sel = _np.asarray(self._mda_selection.indices, dtype=_np.int64)
        wat = _np.asarray(self._water.indices, dtype=_np.int64)
        group_names, ent_group = _np.unique(_np.asarray(self._water_ids, dtype=object).astype(str), return_inverse=True) \
            if len(wat) else (_np.zeros(0, dtype=str), _np.zeros(0, _np.int64))
        n_sel, n_wat = len(sel), len(wat)
        minus = _np.full(n_sel, -1, _np.int32)
        t = dict(n_systems=1,
                 sys_atom_off=_np.asarray([0, self._n_atoms], _np.int64),
                 sys_sel_off=_np.asarray([0, n_sel], _np.int32), sys_wat_off=_np.asarray([0, n_wat], _np.int32),
                 sel_atom=sel.astype(_np.int32), sel_don=minus, sel_acc=minus.copy(), sel_wat=minus.copy(),
                 sel_bb=_np.zeros(n_sel, _np.uint8), wat_atom=wat.astype(_np.int32),
                 wat_ent=_np.arange(n_wat, dtype=_np.int32),
                 ent_resid=_np.asarray(self._cols.resids[wat], dtype=_np.int64).astype(_np.int32),
                 ent_group=_np.asarray(ent_group, _np.int32), ent_residue=_np.asarray(ent_group, _np.int32),
                 ent_h_off=_np.zeros(n_wat + 1, _np.int32), ent_h_atom=_np.zeros(0, _np.int32))
        return t, [str(s) for s in group_names]

    def _pair_tables(self):
        """per_residue=True: selected atoms as entries 0..n_sel-1 (key group = residue id), waters as entries after them."""
        sel = _np.asarray(self._mda_selection.indices, dtype=_np.int64)
        wat = _np.asarray(self._water.indices, dtype=_np.int64)
        if len(_np.intersect1d(sel, wat)):
            raise NotImplementedError('set_presence_around(per_residue=True) with waters of the water group in the selection')
        sel_ids = list(_ids(self._cols, sel, False))
        ids = sel_ids + self._water_ids
        group_names, ent_group = _np.unique(_np.asarray(ids, dtype=object).astype(str), return_inverse=True)
        n_sel, n_wat = len(sel), len(wat)
        minus = _np.full(n_sel, -1, _np.int32)
        resid = _np.concatenate([_np.asarray(self._cols.resids[sel], dtype=_np.int64),
                                 _np.asarray(self._cols.resids[wat], dtype=_np.int64)])
        t = dict(n_systems=1,
                 sys_atom_off=_np.asarray([0, self._n_atoms], _np.int64),
                 sys_sel_off=_np.asarray([0, n_sel], _np.int32), sys_wat_off=_np.asarray([0, n_wat], _np.int32),
                 sel_atom=sel.astype(_np.int32), sel_don=minus, sel_acc=_np.arange(n_sel, dtype=_np.int32), sel_wat=minus.copy(),
                 sel_bb=_np.zeros(n_sel, _np.uint8), wat_atom=wat.astype(_np.int32),
                 wat_ent=(n_sel + _np.arange(n_wat)).astype(_np.int32),
                 ent_resid=resid.astype(_np.int32), ent_group=_np.asarray(ent_group, _np.int32),
                 ent_residue=_np.asarray(ent_group, _np.int32),
                 ent_h_off=_np.zeros(n_sel + n_wat + 1, _np.int32), ent_h_atom=_np.zeros(0, _np.int32))
        is_water_group = _np.zeros(len(group_names), dtype=bool)
        is_water_group[ent_group[n_sel:]] = True
        return t, [str(g) for g in group_names], is_water_group

    def _submit_all(self, ctx):
        block = max(1, min(self.nb_frames, int((256 << 20) // max(1, self._n_atoms * 12))))
        frames = getattr(self._universe.trajectory, "frames", None)
        fi = self._frame_indices
        if hasattr(frames, "xtc_block") and fi.step == 1:
            # XTC: compressed frame records, decoded on the device
            i = fi.start
            while i < fi.stop:
                j = i + frames.xtc_count(i, min(fi.stop, i + block))
                raw, off = frames.xtc_block(i, j)
                ctx.submit_xtc(raw, off, self._n_atoms, j - i, scale=getattr(frames, "LENGTH_SCALE", 10.0))
                i = j
        elif getattr(frames, "planar_block", None) is not None and fi.step == 1:
            # trajectory file records go to the device as they are and are decoded (interleaved) there
            i = fi.start
            while i < fi.stop:
                j = i + frames.planar_count(i, min(fi.stop, i + block))          # (a block ends at a file boundary)
                raw, stride, xo, yo, zo = frames.planar_block(i, j)
                ctx.submit_planar(raw, stride, xo, yo, zo, self._n_atoms, j - i)
                i = j
        else:
            for xyz in _frame_blocks(self._universe, fi, block):
                ctx.submit(xyz)

    def _presence_per_residue(self, around_distance):
        result = {}
        if len(self._water) and self.nb_frames:
            ctx = self._context()
            cached = self.__dict__.get("_packed_pairs")
            if cached is None:
                cached = self._packed_pairs = self._pair_tables()
            tables, group_names, is_water_group = cached
            if getattr(ctx, "_uploaded", None) is not tables:
                ctx.set_topology(tables)
                ctx._uploaded = tables
            r = float(around_distance)
            ctx.set_params(r, r, -1.0, False, False, _native.METHOD_WATER_AROUND)
            ctx.begin(self.nb_frames)
            self._submit_all(ctx)
            keys, words = ctx.results()
            self.last_run_stats = ctx.timers()
            if len(keys):
                bits = _np.unpackbits(words.view(_np.uint8), axis=1, bitorder="little")[:, :self.nb_frames].astype(bool)
                first = (keys >> _np.uint64(32)).astype(_np.int64)
                second = (keys & _np.uint64(0xFFFFFFFF)).astype(_np.int64)
                rows = {}
                for i, (a, b) in enumerate(zip(first, second)):
                    wa, wb = is_water_group[a], is_water_group[b]
                    if wa == wb:                     # water - water pairs of the search are not part of this result
                        continue
                    s_, w_ = (b, a) if wa else (a, b)
                    rows[group_names[s_] + ':' + group_names[w_]] = bits[i]
                order = sorted(rows, key=lambda k: int(rows[k].argmax()))
                result = {k: rows[k] for k in order}
        self.initial_results = result
        self.filtered_results = result

    def _context(self):
        if self._ctx is None:
            self._ctx = _native.Context(0 if self._device is None else self._device)
            for name, value in self._native_options.items():
                self._ctx.set_option(name, value)
        return self._ctx

    def set_presence_around(self, around_distance, per_residue=False):
        """hydration.py:71-107 (per_residue=False)."""
        if per_residue:
            return self._presence_per_residue(around_distance)
        result = {}
        if len(self._water) and self.nb_frames:
            ctx = self._context()
            cached = self.__dict__.get("_packed")
            if cached is None:
                cached = self._packed = self._tables()
            tables, group_names = cached
            if getattr(ctx, "_uploaded", None) is not tables:
                ctx.set_topology(tables)
                ctx._uploaded = tables
            ctx.set_params(1.0, float(around_distance), -1.0, False, False, _native.METHOD_WATER_PRESENCE)
            ctx.begin(self.nb_frames)
            self._submit_all(ctx)
            keys, words = ctx.results()
            self.last_run_stats = ctx.timers()
            if len(keys):
                bits = _np.unpackbits(words.view(_np.uint8), axis=1, bitorder="little")[:, :self.nb_frames].astype(bool)
                rows = {group_names[int(k)]: bits[i] for i, k in enumerate(keys)}
                # the reference's dict grows in order of first appearance: frame, then water index (hydration.py:99-104)
                index_of = {}
                for i, w in enumerate(self._water_ids):
                    index_of.setdefault(w, i)
                order = sorted(rows, key=lambda w: (int(rows[w].argmax()), index_of[w]))
                result = {w: rows[w] for w in order}
        self.initial_results = result
        self.filtered_results = result

    def filter_occupancy(self, min_occupancy, use_filtered=True):     # hydration.py:109-115
        results = self.filtered_results if use_filtered else self.initial_results
        if len(results) == 0: raise AssertionError('nothing to filter!')
        filtered_result = {key: results[key] for key in results if _np.mean(results[key]) > min_occupancy}
        if self.applied_filters['occupancy'] != None:
            self.applied_filters['occupancy'] = max(self.applied_filters['occupancy'], min_occupancy)
        else:
            self.applied_filters['occupancy'] = min_occupancy
        self.filtered_results = filtered_result
