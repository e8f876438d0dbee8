"""Host-side topology compiler: everything the reference constructor derives, as flat arrays.

Mirrors (vectorised, written from scratch) the construction chain of the reference
``HbondAnalysis.__init__`` -> ``NetworkAnalysis.__init__`` -> ``BasicFunctionality.__init__``
(``cgraphs/mdhbond/hbond.py:42-55``, ``network.py:40-97``, ``basic.py:34-69``) and
``helpfunctions.Selection`` / ``MDA_info_list`` (``helpfunctions.py:44-83,107-117``), and flattens the
result into the tables of ``hb_topology`` (``include/hbgpu.h``).
"""
from __future__ import annotations

import numpy as np

# helpfunctions.py:30-34
R_COVALENT = {"N": 1.31, "O": 1.31, "P": 1.58, "S": 1.55}
R_COVALENT_DEFAULT = 1.5
DONOR_NAMES_GLOBAL = {'OH2', 'OW', 'NE', 'NH1', 'NH2', 'ND2', 'SG', 'NE2', 'ND1', 'NZ', 'OG', 'OG1', 'NE1', 'OH',
                      'OE1', 'OE2', 'N16', 'OD1', 'OD2'}
ACCEPTOR_NAMES_GLOBAL = {'OH2', 'OW', 'OD1', 'OD2', 'SG', 'OE1', 'OE2', 'ND1', 'NE2', 'SD', 'OG', 'OG1', 'OH'}
WATER_DEFINITION = '(resname TIP3 and name OH2) or (resname HOH and name O) or (resname TIP4 and name OH2)'


class AtomColumns:
    """names / resnames / resids / segids / types / resindices of all atoms of a universe."""

    def __init__(self, universe):
        t = getattr(universe, "topology", None)
        if t is not None and hasattr(t, "resindices") and hasattr(t, "names"):   # cgraphs_b200.universe.Universe
            self.names, self.resnames, self.resids = t.names, t.resnames, t.resids
            self.segids, self.types, self.resindices = t.segids, t.types, t.resindices
            self._from_arrays()
            return
        atoms = universe.atoms
        # real MDAnalysis
        self.names = np.asarray(atoms.names, dtype=object)
        self.resnames = np.asarray(atoms.resnames, dtype=object)
        self.resids = np.asarray(atoms.resids, dtype=np.int64)
        self.segids = np.asarray(atoms.segids, dtype=object)
        try:
            self.types = np.asarray(atoms.types, dtype=object)
        except Exception:
            self.types = np.asarray([""] * len(self.names), dtype=object)
        self.resindices = np.asarray(atoms.resindices, dtype=np.int64)
        self._from_arrays()

    def _from_arrays(self):
        self.n_atoms = len(self.names)
        names = self.names
        self.first = np.fromiter((n[0] if len(n) else "" for n in names), dtype="U1", count=self.n_atoms)
        self.name_len = np.fromiter((len(n) for n in names), dtype=np.int64, count=self.n_atoms)
        two = np.fromiter((n[:2] for n in names), dtype="U2", count=self.n_atoms)
        # helpfunctions.py:65 - what counts as a hydrogen when looking for bonded hydrogens
        self.h_like = (self.first == "H") | np.isin(two, ["1H", "2H", "3H"]) | (self.types == "H")
        # residue -> atoms (ascending atom index, which is the order of `residue.atoms`)
        order = np.argsort(self.resindices, kind="stable")
        sr = self.resindices[order]
        n_res = int(sr[-1]) + 1 if self.n_atoms else 0
        self.res_order = order
        self.res_start = np.searchsorted(sr, np.arange(n_res + 1))


def _expand_ranges(starts, counts):
    """Concatenate ranges [starts[i], starts[i]+counts[i]) -> (flat indices, owner index)."""
    counts = np.asarray(counts, dtype=np.int64)
    total = int(counts.sum())
    owner = np.repeat(np.arange(len(counts), dtype=np.int64), counts)
    if total == 0:
        return np.zeros(0, np.int64), owner
    offs = np.cumsum(counts) - counts
    flat = np.arange(total, dtype=np.int64) - np.repeat(offs, counts) + np.repeat(np.asarray(starts, np.int64), counts)
    return flat, owner


def _ids(cols, atoms, detailed):
    """helpfunctions.py:107-117 (MDA_info_list)."""
    seg, rn, ri, nm = cols.segids, cols.resnames, cols.resids, cols.names
    if detailed:
        return [str(seg[i]) + '-' + str(rn[i]) + '-' + str(int(ri[i])) + '-' + nm[i] for i in atoms]
    return [str(seg[i]) + '-' + str(rn[i]) + '-' + str(int(ri[i])) for i in atoms]


class CompiledTopology:
    """Flat tables for one system (see include/hbgpu.h for the meaning of each array)."""

    def __init__(self, cols, sel_atoms, water_atoms, pos0, *, additional_donors=(), additional_acceptors=(),
                 exclude_donors=(), exclude_acceptors=(), check_angle=True, residuewise=True,
                 add_donors_without_hydrogen=False):
        self.donor_names = DONOR_NAMES_GLOBAL.union(additional_donors) - set(exclude_donors)
        self.acceptor_names = ACCEPTOR_NAMES_GLOBAL.union(additional_acceptors) - set(exclude_acceptors)
        self.residuewise = residuewise
        self.check_angle = check_angle
        sel = np.asarray(sel_atoms, dtype=np.int64)
        water_atoms = np.asarray(water_atoms, dtype=np.int64)
        names = cols.names[sel]
        polar = (cols.name_len[sel] > 1) & np.isin(cols.first[sel], ["N", "O"])        # helpfunctions.py:57,74
        don_cand = np.isin(names, list(self.donor_names)) | polar
        acc_mask = np.isin(names, list(self.acceptor_names)) | polar

        if add_donors_without_hydrogen:
            donors = sel[don_cand]
            sel_h = np.zeros(0, np.int64)
            don_h_cnt = np.zeros(len(donors), np.int64)
        else:
            cand = sel[don_cand]
            # hydrogen-like atoms of each residue, in residue order
            hl_sorted = cols.h_like[cols.res_order]
            hl_csum = np.concatenate([[0], np.cumsum(hl_sorted)])
            h_atoms_by_res = cols.res_order[hl_sorted]                    # CSR values
            h_res_start = hl_csum[cols.res_start]                         # CSR offsets per residue
            r = cols.resindices[cand]
            flat, owner = _expand_ranges(h_res_start[r], h_res_start[r + 1] - h_res_start[r])
            o = h_atoms_by_res[flat]
            a = cand[owner]
            pa, po = pos0[a], pos0[o]                                      # float32
            d = pa - po
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            dist = np.sqrt(d2)                                             # float32, helpfunctions.py:65
            rc = np.array([R_COVALENT.get(f, R_COVALENT_DEFAULT) for f in cols.first[a]], dtype=np.float32) \
                if len(a) else np.zeros(0, np.float32)
            bonded = dist < rc
            cnt = np.bincount(owner[bonded], minlength=len(cand)).astype(np.int64)
            keep = cnt > 0
            donors = cand[keep]
            sel_h = o[bonded]                      # already grouped by candidate in selection order
            don_h_cnt = cnt[keep]
        acceptors = sel[acc_mask]
        if len(donors) + len(acceptors) == 0:
            raise AssertionError('neither donors nor acceptors in the selection')
        nD, nA, nW = len(donors), len(acceptors), len(water_atoms)
        self.nD, self.nA, self.nW = nD, nA, nW
        self.donors, self.acceptors, self.water_atoms = donors, acceptors, water_atoms
        self.first_water = nD + nA
        self.entry_atom = np.concatenate([donors, acceptors, water_atoms]).astype(np.int64)
        nE = len(self.entry_atom)

        # water hydrogens: residue.atoms[1:] of every water, concatenated (network.py:69-77)
        if add_donors_without_hydrogen or nW == 0:
            water_h = np.zeros(0, np.int64)
        else:
            r = cols.resindices[water_atoms]
            flat, _ = _expand_ranges(cols.res_start[r] + 1, cols.res_start[r + 1] - cols.res_start[r] - 1)
            water_h = cols.res_order[flat]
        if len(sel_h) == 0 and len(water_h) == 0 and check_angle:
            raise AssertionError('There are no possible hbond donors in the selection and no water. Since '
                                 'check_angle is True, hydrogen is needed for the calculations!')
        self.h_atom = np.concatenate([sel_h, water_h]).astype(np.int64)        # self._hydrogen
        self.first_water_h = len(sel_h)
        # heavy2hydrogen as CSR over entries (network.py:86): donors -> their ranges, acceptors -> none,
        # water k -> hydrogens [2k, 2k+1] of the concatenated water-hydrogen list
        n_wlists = (len(water_h) + 1) // 2
        h_cnt = np.zeros(nE, np.int64)
        h_cnt[:nD] = don_h_cnt
        self.water_h_lists = n_wlists
        wl = min(n_wlists, nW)
        h_cnt[self.first_water:self.first_water + wl] = 2
        if len(water_h) % 2 == 1 and wl == n_wlists and wl > 0:
            h_cnt[self.first_water + wl - 1] = 1       # the reference would index past the end here
            self.water_h_ragged = True
        else:
            self.water_h_ragged = False
        self.h_off = np.concatenate([[0], np.cumsum(h_cnt)]).astype(np.int64)
        don_flat = np.arange(len(sel_h), dtype=np.int64)
        wat_flat = len(sel_h) + np.arange(int(h_cnt[self.first_water:].sum()), dtype=np.int64)
        self.h_idx = np.concatenate([don_flat, wat_flat])                         # into h_atom
        assert len(self.h_idx) == self.h_off[-1]

        # ids (network.py:65-66,87-89)
        sel_entries = self.entry_atom[:self.first_water]
        self.water_ids = _ids(cols, water_atoms, False)
        self.water_ids_atomwise = _ids(cols, water_atoms, True)
        self.all_ids = _ids(cols, sel_entries, not residuewise) + self.water_ids
        self.all_ids_atomwise = _ids(cols, sel_entries, True) + self.water_ids_atomwise
        self.resids = np.array([int(s.split('-')[2]) for s in self.all_ids], dtype=np.int64)
        self.backbone = np.array([(s.split('-')[3] in ('O', 'N')) for s in self.all_ids_atomwise], dtype=bool)

        # points
        sel_pts = np.union1d(donors, acceptors)
        self.sel_atom = sel_pts
        self.sel_don = np.full(len(sel_pts), -1, np.int64)
        self.sel_acc = np.full(len(sel_pts), -1, np.int64)
        self.sel_wat = np.full(len(sel_pts), -1, np.int64)
        self.sel_don[np.searchsorted(sel_pts, donors)] = np.arange(nD)
        self.sel_acc[np.searchsorted(sel_pts, acceptors)] = nD + np.arange(nA)
        in_sel = np.isin(water_atoms, sel_pts)
        w_in = np.nonzero(in_sel)[0]
        self.sel_wat[np.searchsorted(sel_pts, water_atoms[w_in])] = self.first_water + w_in
        w_out = np.nonzero(~in_sel)[0]
        self.wat_atom = water_atoms[w_out]
        self.wat_ent = self.first_water + w_out
        ent_of_pt = np.where(self.sel_don >= 0, self.sel_don, self.sel_acc)
        self.sel_bb = self.backbone[ent_of_pt].astype(np.uint8)

    # -- key groups --------------------------------------------------------------------------------
    def key_ids(self, method):
        """The id list the reference concatenates into bond keys for `method`
        (hbond.py:122 always `_all_ids`; hbond.py:276-277 `_all_ids` / `_all_ids_atomwise`)."""
        if method == "in_selection" or self.residuewise:
            return self.all_ids
        return self.all_ids_atomwise

    def heavy2hydrogen(self):
        """The reference's list-of-lists form (network.py:86), for introspection and tests."""
        out = []
        for e in range(len(self.entry_atom)):
            out.append(list(self.h_idx[self.h_off[e]:self.h_off[e + 1]]))
        return out


def pack_systems(tops, methods_ids, n_atoms_list, extra_ids=None):
    """Concatenate CompiledTopology objects of several systems into the hb_topology arrays.

    `methods_ids[i]` is the id list (one string per entry) that forms the keys of system i.  Key group
    ids index the sorted list of distinct id strings over all systems, returned as `group_names`.
    `extra_ids` (e.g. ion ids) join the same id table; their group ids are returned as a third value.
    """
    all_ids = [s for ids in methods_ids for s in ids]
    n_ent_ids = len(all_ids)
    if extra_ids is not None:
        all_ids = all_ids + list(extra_ids)
    group_names, ent_group = np.unique(np.asarray(all_ids, dtype=object).astype(str), return_inverse=True) \
        if all_ids else (np.zeros(0, dtype=str), np.zeros(0, np.int64))
    extra_group = np.asarray(ent_group[n_ent_ids:], np.int32)
    ent_group = ent_group[:n_ent_ids]
    all_ids = all_ids[:n_ent_ids]
    res3 = ['-'.join(s.split('-')[:3]) for s in all_ids]
    _, ent_residue = np.unique(np.asarray(res3, dtype=object).astype(str), return_inverse=True) \
        if res3 else (None, np.zeros(0, np.int64))
    t = dict(n_systems=len(tops))
    atom_off, sel_off, wat_off = [0], [0], [0]
    ent_base = 0
    cols = {k: [] for k in ("sel_atom", "sel_don", "sel_acc", "sel_wat", "sel_bb", "wat_atom", "wat_ent",
                            "ent_resid", "ent_h_cnt", "ent_h_atom")}
    for top, na in zip(tops, n_atoms_list):
        shift = lambda a: np.where(a >= 0, a + ent_base, -1)
        cols["sel_atom"].append(top.sel_atom)
        cols["sel_don"].append(shift(top.sel_don))
        cols["sel_acc"].append(shift(top.sel_acc))
        cols["sel_wat"].append(shift(top.sel_wat))
        cols["sel_bb"].append(top.sel_bb)
        cols["wat_atom"].append(top.wat_atom)
        cols["wat_ent"].append(top.wat_ent + ent_base)
        cols["ent_resid"].append(top.resids)
        cols["ent_h_cnt"].append(np.diff(top.h_off))
        cols["ent_h_atom"].append(top.h_atom[top.h_idx])
        atom_off.append(atom_off[-1] + int(na))
        sel_off.append(sel_off[-1] + len(top.sel_atom))
        wat_off.append(wat_off[-1] + len(top.wat_atom))
        ent_base += len(top.entry_atom)
    cat = lambda k, dt: np.concatenate(cols[k]).astype(dt) if cols[k] else np.zeros(0, dt)
    t["sys_atom_off"] = np.asarray(atom_off, np.int64)
    t["sys_sel_off"] = np.asarray(sel_off, np.int32)
    t["sys_wat_off"] = np.asarray(wat_off, np.int32)
    for k in ("sel_atom", "sel_don", "sel_acc", "sel_wat", "wat_atom", "wat_ent", "ent_h_atom"):
        t[k] = cat(k, np.int32)
    t["sel_bb"] = cat("sel_bb", np.uint8)
    resid = cat("ent_resid", np.int64)
    if len(resid) and (resid.max() > 2**31 - 1 or resid.min() < -2**31):
        raise ValueError("resid outside int32")
    t["ent_resid"] = resid.astype(np.int32)
    t["ent_group"] = np.asarray(ent_group, np.int32)
    t["ent_residue"] = np.asarray(ent_residue, np.int32)
    t["ent_h_off"] = np.concatenate([[0], np.cumsum(cat("ent_h_cnt", np.int64))]).astype(np.int32)
    if extra_ids is not None:
        return t, [str(s) for s in group_names], extra_group
    return t, [str(s) for s in group_names]
