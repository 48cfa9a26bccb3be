"""Host side of the hydrophobic-PMF term: tables, atom typing, bond
connectivity and the ctypes binding of ``mcsce_hpmf_*`` / ``mcsce_sasa``.

Reference: ``libs/libhpmf.py`` (three-Gaussian PMF over exposed carbon pairs), ``libs/libsasa.py``
(Hasel-Hendrickson-Still approximate surface), ``core/definitions.py:181-511`` (atom types).
The reference's own wiring of the term raises (``libs/libenergy.py:110-113`` vs ``:217``); the two
closures it is made of run when called directly and are what the CUDA kernels reproduce, for batches
of whole structures.  No CPU fallback: the calculators need the CUDA library and a device.
"""
from __future__ import annotations

import ctypes as C
import gzip
import json
from functools import lru_cache
from pathlib import Path

import numpy as np

from mcsce_b200._lib import SASA_MAX_BONDS, HpmfDesc, check, load_library
from mcsce_b200.chem.tables import INTER_BOND_1, load_tables


@lru_cache(maxsize=1)
def load_sasa_tables(path=None):
    """P/R parameters (libsasa.py:21-74), residue/atom -> type map (definitions.py:181-511) and the
    Gaussian constants (libhpmf.py:17-20); packed by ``tests/golden/make_golden_hpmf.py``."""
    path = Path(path or Path(__file__).resolve().parent / "data" / "sasa_tables.json.gz")
    with gzip.open(path, "rb") as f:
        return json.loads(f.read().decode())


def convert_atom_types(atom_labels, residue_numbers, residue_labels):
    """Same contract as ``libhpmf.convert_atom_types`` (libhpmf.py:22-68): boolean filter of the atoms
    that carry a SASA type (carbon-bound hydrogens do not) and the list of their types.  Residue 1 gets
    the charged N-terminus types, ``OXT`` is ``O-``; a residue without a type table raises ``KeyError``
    as in the reference (its table covers the 20 standard residues with HID/HIE/HIP)."""
    type_map = load_sasa_tables()["atom_types"]
    atom_filter, new_atom_types = [], []
    for label, resnum, resname in zip(atom_labels, residue_numbers, residue_labels):
        label, resname = str(label), str(resname)
        if resnum == 1 and label == "N":
            new_type = "N-SP3"
        elif resnum == 1 and label in ("H1", "H2", "H3"):
            new_type = "H-NH+"
        elif label == "OXT":
            new_type = "O-"
        else:
            new_type = type_map[resname][label]
        atom_filter.append(new_type is not None)
        if new_type is not None:
            new_atom_types.append(new_type)
    return np.array(atom_filter), new_atom_types


def bond_pairs(atom_labels, residue_numbers, residue_labels, tables=None):
    """Covalent bonds (i < j, indices into the given atom list) under the reference's rule-based
    1-bond mask: template topology inside a residue, C(n)-N(n+1) between residues
    (libenergy.py:97-103, build_definitions.py:460-462) - the matrix the commented-out lines at
    libenergy.py:110-113 would have built."""
    t = tables or load_tables()
    by_res = {}
    for idx, (a, n, r) in enumerate(zip(atom_labels, residue_numbers, residue_labels)):
        by_res.setdefault(int(n), (str(r), {}))[1][str(a)] = idx
    pairs = set()
    for n, (label, atoms) in by_res.items():
        nb = t.bonds[label]
        for a, i in atoms.items():
            for b in nb.get(a, ()):
                j = atoms.get(b)
                if j is not None and j != i:
                    pairs.add((min(i, j), max(i, j)))
        nxt = by_res.get(n + 1)
        if nxt is not None:
            for a, partners in INTER_BOND_1.items():
                for b in partners:
                    if a in atoms and b in nxt[1]:
                        i, j = atoms[a], nxt[1][b]
                        pairs.add((min(i, j), max(i, j)))
    out = np.array(sorted(pairs), dtype=np.int64).reshape(-1, 2)
    return out[:, 0], out[:, 1]


def _bonded_table(n_typed, bond_connectivities):
    """(n_typed, SASA_MAX_BONDS) int32 neighbour table from an N x N matrix or an (i, j) pair list."""
    if isinstance(bond_connectivities, tuple):
        bi, bj = (np.asarray(x, dtype=np.int64) for x in bond_connectivities)
    else:
        m = np.asarray(bond_connectivities)
        if m.shape != (n_typed, n_typed):
            raise ValueError("bond_connectivities must be %d x %d, got %s" % (n_typed, n_typed, m.shape))
        bi, bj = np.nonzero(m)
    table = np.full((n_typed, SASA_MAX_BONDS), -1, dtype=np.int32)
    fill = np.zeros(n_typed, dtype=np.int64)
    seen = set()
    for i, j in zip(bi.tolist(), bj.tolist()):
        for a, b in ((i, j), (j, i)):
            if a == b or (a, b) in seen:
                continue
            seen.add((a, b))
            if fill[a] >= SASA_MAX_BONDS:
                raise ValueError("atom %d has more than %d bonded neighbours with a SASA type" % (a, SASA_MAX_BONDS))
            table[a, fill[a]] = b
            fill[a] += 1
    return table


class HpmfCalculator:
    """Device calculator of SA_i and V_hpmf for batches of structures (one handle, one device)."""

    def __init__(self, sasa_atom_types, atom_filter, bond_connectivities, threshold=None, r_solvent=None, device=0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("mcsce_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = load_library()
        tb = load_sasa_tables()
        self.device = torch.device("cuda", int(device))
        atom_filter = np.asarray(atom_filter, dtype=bool)
        self.n_atoms = int(len(atom_filter))
        self.typed_atom = np.ascontiguousarray(np.nonzero(atom_filter)[0], dtype=np.int32)
        self.n_typed = int(len(self.typed_atom))
        if len(sasa_atom_types) != self.n_typed:
            raise ValueError("%d atom types for %d selected atoms" % (len(sasa_atom_types), self.n_typed))
        names = sorted(tb["R"])
        index = {n: k for k, n in enumerate(names)}
        self.type_of = np.ascontiguousarray([index[str(t)] for t in sasa_atom_types], dtype=np.int32)   # KeyError: unknown type
        self.bonded = np.ascontiguousarray(_bonded_table(self.n_typed, bond_connectivities))
        self.type_R = np.array([tb["R"][n] for n in names], dtype=np.float64)
        self.type_P = np.array([tb["P"][n] for n in names], dtype=np.float64)
        self.type_is_carbon = np.array([n.startswith("C") for n in names], dtype=np.uint8)
        self.carbon = self.type_is_carbon[self.type_of].astype(bool)
        d = HpmfDesc()
        d.n_atoms, d.n_typed, d.n_types = self.n_atoms, self.n_typed, len(names)
        d.typed_atom = self.typed_atom.ctypes.data_as(C.POINTER(C.c_int32))
        d.type_of = self.type_of.ctypes.data_as(C.POINTER(C.c_int32))
        d.bonded = self.bonded.ctypes.data_as(C.POINTER(C.c_int32))
        d.type_R = self.type_R.ctypes.data_as(C.POINTER(C.c_double))
        d.type_P = self.type_P.ctypes.data_as(C.POINTER(C.c_double))
        d.type_is_carbon = self.type_is_carbon.ctypes.data_as(C.POINTER(C.c_uint8))
        d.r_solvent = float(tb["r_solvent"] if r_solvent is None else r_solvent)
        d.pij_bonded, d.pij_non_bonded = float(tb["pij_bonded"]), float(tb["pij_non_bonded"])
        for k in range(3):
            d.hpmf_c[k], d.hpmf_w[k], d.hpmf_h[k] = (float(tb["hpmf"][x][k]) for x in ("C", "W", "H"))
        d.sa_threshold = float(tb["hpmf"]["AC"] if threshold is None else threshold)
        h = C.c_void_p()
        check(self.lib.mcsce_hpmf_create(C.byref(h), int(device), C.byref(d)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mcsce_hpmf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _device_coords(self, coords):
        torch = self.torch
        if not torch.is_tensor(coords):
            coords = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float32))
        x = coords.to(device=self.device, dtype=torch.float32).contiguous()
        single = x.dim() == 2
        if single:
            x = x[None]
        if x.dim() != 3 or x.shape[1] != self.n_atoms or x.shape[2] != 3:
            raise ValueError("coords must be (%d, 3) or (B, %d, 3), got %s" % (self.n_atoms, self.n_atoms, tuple(coords.shape)))
        return x, single

    def sasa(self, coords):
        """(B, n_atoms, 3) or (n_atoms, 3) float32 -> (B, n_typed) / (n_typed,) float64 device tensor."""
        torch = self.torch
        x, single = self._device_coords(coords)
        out = torch.empty((x.shape[0], self.n_typed), dtype=torch.float64, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.mcsce_sasa(self.h, int(x.shape[0]), x.data_ptr(), out.data_ptr(), st))
        return out[0] if single else out

    def energy(self, coords, return_sasa=False):
        """V_hpmf of every structure: (B,) float64 device tensor (0-d for a single structure)."""
        torch = self.torch
        x, single = self._device_coords(coords)
        sasa = torch.empty((x.shape[0], self.n_typed), dtype=torch.float64, device=self.device)
        v = torch.empty(x.shape[0], dtype=torch.float64, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.mcsce_hpmf_energy(self.h, int(x.shape[0]), x.data_ptr(), sasa.data_ptr(), v.data_ptr(), st))
        if single:
            v, sasa = v[0], sasa[0]
        return (v, sasa) if return_sasa else v

    def trial_energies(self, coords, first_new, n_new, trial_xyz, mask=None):
        """V_hpmf of (atoms before ``first_new`` + each trial's ``n_new`` new atoms): ``coords`` (C, n_atoms, 3),
        ``trial_xyz`` (C, T, n_new, 3), ``mask`` (C, T) bool or None -> (C, T) float64 device tensor, 0 where masked.
        The environment's surface areas and pair fields are evaluated once per conformer; every trial only adds its
        own factors and pair terms (``mcsce_hpmf_trials``)."""
        torch = self.torch
        x = coords.to(device=self.device, dtype=torch.float32).contiguous()
        t = trial_xyz.to(device=self.device, dtype=torch.float32).contiguous()
        C_, T = int(t.shape[0]), int(t.shape[1])
        assert x.shape == (C_, self.n_atoms, 3) and t.shape[2] == n_new and t.shape[3] == 3
        out = torch.empty((C_, T), dtype=torch.float64, device=self.device)
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            assert m.shape == (C_, T)
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.mcsce_hpmf_trials(self.h, C_, x.data_ptr(), int(first_new), int(n_new), T, t.data_ptr(),
                                         m.data_ptr() if m is not None else None, out.data_ptr(), st))
        return out

    def carry(self, enable):
        """Let consecutive ``trial_energies`` calls of one growth (same ``coords`` buffer, rising ``first_new``) extend
        the environment's state instead of re-deriving it; every call of this method forgets the kept state."""
        check(self.lib.mcsce_hpmf_carry(self.h, 1 if enable else 0))

    def last_ms(self):
        ms = (C.c_double * 2)()
        check(self.lib.mcsce_hpmf_last_ms(self.h, ms))
        return float(ms[0]), float(ms[1])


def calculator_for_labels(atom_labels, residue_numbers, residue_labels, device=0, tables=None):
    """Calculator for a labelled atom list (any order): types by ``convert_atom_types``, bonds by
    ``bond_pairs`` restricted to the typed atoms."""
    atom_filter, types = convert_atom_types(atom_labels, residue_numbers, residue_labels)
    bi, bj = bond_pairs(atom_labels, residue_numbers, residue_labels, tables)
    typed_index = np.cumsum(atom_filter) - 1
    keep = atom_filter[bi] & atom_filter[bj]
    return HpmfCalculator(types, atom_filter, (typed_index[bi[keep]], typed_index[bj[keep]]), device=device)


def calculator_for_structure(structure, device=0):
    return calculator_for_labels(structure.atom_labels, structure.res_nums, structure.res_labels, device=device)


def growth_term(engine, calc=None, max_bytes=1 << 30, brute_force=False, carry=None):
    """``extra_term`` for :meth:`GrowthEngine.grow_stepwise`: V_hpmf of every partially grown structure a trial
    would produce - the input atoms, the side chains placed so far and the trial's - which is what the reference's
    coordinate-based energy terms receive (libenergy.py:806-808 with the level's atom list, side_chain_builder.py:
    93-104), evaluated by ``mcsce_hpmf_energy`` on structures whose later atoms are parked (absent).

    Default: ``mcsce_hpmf_trials`` - the environment once per conformer and residue, O(N^2), then every trial adds
    its own factors and pair terms, O(n_new N).  ``brute_force=True``: one whole-structure evaluation per
    (conformer, trial), O(T N^2) per residue and conformer; same numbers, kept as the check of the first.
    ``carry`` (default: on unless ``MCSCE_B200_HPMF_CARRY=0``): the environment's areas, weights and fields are kept
    from residue to residue and extended by the atoms placed since (``mcsce_hpmf_carry``) instead of re-derived."""
    import os
    if carry is None:
        carry = os.environ.get("MCSCE_B200_HPMF_CARRY", "1") != "0"
    torch = engine.torch
    sysm = engine.system
    if calc is None:
        calc = calculator_for_labels([str(a) for a in sysm.names], [int(r) for r in sysm.resnums],
                                     [str(r) for r in sysm.resnames], device=engine.device_index)
    N = sysm.n_atoms
    per = max(1, int(max_bytes // (N * 12)))
    parked = 1.0e15

    def term(st, coords, xyz, finite):
        if st is None:                                     # level 0: the input atoms alone, the same for every conformer
            calc.carry(carry and not brute_force)          # a new growth: forget the previous one's environment
            s0 = coords[:1].clone()
            s0[:, sysm.n_input:] = parked
            return calc.energy(s0).expand(coords.shape[0])
        if not brute_force:
            return calc.trial_energies(coords, st.sc_start, st.n_sc, xyz, finite)
        out = torch.zeros(finite.shape, dtype=torch.float64, device=coords.device)
        idx = finite.nonzero()
        end = st.sc_start + st.n_sc
        for lo in range(0, idx.shape[0], per):
            c, t = idx[lo:lo + per, 0], idx[lo:lo + per, 1]
            s = coords[c]
            s[:, st.sc_start:end] = xyz[c, t]
            s[:, end:] = parked
            out[c, t] = calc.energy(s)
        return out

    term.calculator = calc
    return term
