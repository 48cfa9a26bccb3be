"""Drop-in for ``mcsce.core.side_chain_builder`` (side_chain_builder.py:42-352): same entry points,
the growth loop runs on the GPU for all conformers of a call at once.

* ``initialize_func_calc``          builds the O(N) host tables and uploads them (replaces the L+1
                                    placeholder structures / energy closures, :42-107);
* ``create_side_chain_structure``   one conformer (:109-205);
* ``create_side_chain``             best-of-N / first-valid (:207-281);
* ``create_side_chain_ensemble``    N conformers + ``<n>.pdb`` + ``energies.csv`` (:284-352).

``parallel_worker`` only sized the reference's process pool; it is accepted and ignored (one process
per GPU here; under ``torchrun`` the conformers of a call are sharded over ranks and the results
gathered to rank 0 with one collective, see ``mcsce_b200.sharding``).  The reference's in-place drift
of the cached library rows (quirk Q1, :173) is not reproduced: every conformer perturbs the pristine
library rows, on the device, with its own Philox stream.
"""
from __future__ import annotations

import os
from functools import partial

import numpy as np

from mcsce_b200 import sharding
from mcsce_b200.chem.tables import load_tables
from mcsce_b200.core.rotamer_library import DunbrakRotamerLibrary, ptmRotamerLib
from mcsce_b200.structure import Structure
from mcsce_b200.system import build_system

_rotamer_library = None
_ptm_rotamer_lib = None

#: conformers grown per device launch sequence; larger ensembles are processed in chunks of this size
MAX_CONFORMERS_PER_LAUNCH = 4096
#: the same for the step-by-step growth that carries the ``hpmf`` term
MAX_CONFORMERS_PER_STEPWISE_CALL_UNUSED = 1024

# kept for API compatibility: level 0 = the input structure, level i+1 = with residue i's side chain
sidechain_placeholders = []
energy_calculators = []
_state = {}


def _libraries():
    global _rotamer_library, _ptm_rotamer_lib
    if _rotamer_library is None:
        _rotamer_library = DunbrakRotamerLibrary()
        _ptm_rotamer_lib = ptmRotamerLib()
    return _rotamer_library, _ptm_rotamer_lib


class _LazyCalculators:
    """``energy_calculators[level]`` on demand: the reference prepares every closure up front (63 s at 76
    residues); the fused growth never needs them, so they are only built when a caller indexes the list."""

    def __init__(self, creator, placeholders, kinds):
        self.creator, self.placeholders, self.kinds, self.cache = creator, placeholders, kinds, {}

    def __len__(self):
        return len(self.placeholders)

    def __getitem__(self, level):
        if level < 0:
            level += len(self)
        if level in self.cache:
            return self.cache[level]
        if level > 0 and self.kinds[level - 1] != "grow":
            return None
        s = self.placeholders[level]
        n_terms, c_terms = s.get_terminal_res_atom_arr()
        kw = {}
        if level > 0:
            n_sc = len(s) - len(self.placeholders[level - 1])
            idx = np.arange(len(s))
            kw["partial_indices"] = [idx[-n_sc:], idx[:-n_sc]]
        self.cache[level] = self.creator(s.atom_labels, s.res_nums, s.res_labels, n_terms, c_terms, **kw)
        return self.cache[level]


def _terms_of(efunc_creator):
    kw = getattr(efunc_creator, "keywords", None) or {}
    terms = kw.get("terms")
    if terms is None:
        terms = ["lj", "clash", "coulomb"]
    return tuple(terms)


def initialize_func_calc(efunc_creator, aa_seq=None, structure=None, retain_idxs=[]):
    """Prepare the growth of one sequence.  ``efunc_creator`` is the reference's
    ``partial(prepare_energy_function, batch_size=..., forcefield=..., terms=[...])``; only its ``terms``
    are needed here."""
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.libs.libenergy import _device

    global sidechain_placeholders, energy_calculators
    print("Start preparing energy calculators at different sidechain completion levels")
    if structure is None:
        from mcsce_b200.structure import AA3TO1
        structure = Structure(fasta="".join(AA3TO1.get(r, "X") for r in aa_seq)).build()
    structure = structure.copy()
    if aa_seq is not None and structure is not None:
        # the IDPConfGen bridge passes both: the sequence overrides the residue labels (mcsce_sidechain.py:69-70)
        nums = structure.residues
        for num, res in zip(nums, aa_seq):
            structure.cols["resname"][structure.cols["resseq"] == num] = res
    lib, ptm = _libraries()
    t = load_tables()
    terms = _terms_of(efunc_creator)
    from mcsce_b200.libs.libenergy import _SUPPORTED
    bad = [x for x in terms if x not in _SUPPORTED]
    if bad:
        raise NotImplementedError("energy terms %s are not runnable in the reference either; supported: %s"
                                  % (bad, list(_SUPPORTED)))
    sysm = build_system(structure, retain_idxs=tuple(retain_idxs), terms=terms, rotamer_library=lib, ptm_library=ptm, tables=t)
    eng = _state.get("engine")
    if eng is None:
        eng = GrowthEngine(_device())
    eng.set_system(sysm, t)
    # placeholders: atom bookkeeping per completion level (labels only; coordinates are filled at growth)
    placeholders = [structure.copy()]
    work = structure.copy()
    for st in sysm.steps:
        if st.kind != "skip":
            tm = t.templates[st.resname]
            work.append_atoms([tm.names[k] for k in tm.sc_idx], tm.resname, str(sysm.chains[st.sc_start]), st.resnum,
                              np.zeros((tm.n_sc, 3), np.float32), elements=[tm.elements[k] for k in tm.sc_idx])
        placeholders.append(work.copy())
    sidechain_placeholders = placeholders
    energy_calculators = _LazyCalculators(efunc_creator, placeholders, [st.kind for st in sysm.steps])
    _state.pop("final", None)
    _state.update(engine=eng, system=sysm, structure=structure, retain=tuple(retain_idxs), backbone=None, seed=_state.get("seed", 0))
    _state["hpmf"] = None
    if "hpmf" in terms:
        # coordinate-based term (libenergy.py:213-220, 806-808): evaluated inside the fused loop (mcsce_grow_opts.hpmf)
        from mcsce_b200.hpmf import calculator_for_labels
        _state["hpmf"] = calculator_for_labels([str(a) for a in sysm.names], [int(r) for r in sysm.resnums],
                                               [str(r) for r in sysm.resnames], device=eng.device_index)
    print("Finished preparing all energy functions. Now start conformer generation")


def set_seed(seed):
    """Seed of the on-device Philox streams (the reference draws from NumPy's / numba's global generators)."""
    _state["seed"] = int(seed)
    _state["next_conf"] = 0


def _grow(backbone_coords, beta, n_conf):
    """Grow ``n_conf`` conformers of the prepared sequence on this rank's GPU; returns host arrays."""
    assert len(sidechain_placeholders) > 0, "Energy functions have not yet initialized!"
    eng, sysm = _state["engine"], _state["system"]
    backbone_coords = np.asarray(backbone_coords, dtype=np.float32)
    assert sysm.n_input == backbone_coords.shape[0], \
        "Input structures contain extra sidechain atoms or miss sidechain atoms if fix argument is used"
    key = backbone_coords.tobytes()
    if _state.get("backbone") != key:
        eng.set_backbone(backbone_coords)
        _state["backbone"] = key
    lo, hi = sharding.my_range(n_conf)
    first = _state.get("next_conf", 0)
    _state["next_conf"] = first + n_conf
    single = sharding.world()[1] == 1
    parts = []
    # one rank: results straight into pinned host memory; several ranks: keep them on the device for the gather.
    # Big requests are grown in chunks so that the device workspace and the pinned staging stay bounded.  Conformer
    # streams (Philox) are keyed by the global conformer id, so every conformer sees the same trial rows and uniforms whatever
    # the chunking or sharding; its energies agree to the float32 grouping of the pair sums, which follows the launch geometry
    # (cluster size of the persistent kernel, environment split of the per-kernel path).
    hpmf = _state.get("hpmf")
    for c0 in range(lo, hi, MAX_CONFORMERS_PER_LAUNCH):
        n = min(MAX_CONFORMERS_PER_LAUNCH, hi - c0)
        out = eng.grow(n, beta, seed=_state.get("seed", 0), conf_id0=first + c0, host_out=single, hpmf=hpmf)
        part = (out["coords"], out["energy"], out["ok"])
        parts.append(tuple(x.copy() for x in part) if single else tuple(x.clone() for x in part))
    if not parts:
        local = (np.zeros((0, sysm.n_atoms, 3), np.float32), np.zeros(0), np.zeros(0, np.uint8))
    elif len(parts) == 1:
        local = parts[0]
    elif single:
        local = tuple(np.concatenate([p[k] for p in parts]) for k in range(3))
    else:
        import torch
        local = tuple(torch.cat([p[k] for p in parts]) for k in range(3))
    coords, energy, ok = sharding.gather_results(*local, n_total=n_conf)
    if ok is not None and (ok == 2).any():
        # ok = 2: the hydrophobic PMF gave a conformer up because a neighbour-list capacity was exceeded (mcsce_b200.h)
        import warnings
        warnings.warn("mcsce_b200: %d conformers were given up by the hydrophobic PMF term (capacity exceeded); they are "
                      "reported as failed" % int((ok == 2).sum()))
        ok = np.where(ok == 2, 0, ok).astype(np.uint8)
    return coords, energy, ok


def _final_template():
    """Finished-structure table in output order (residue-number order, libstructure.py:552-555), the
    permutation from growth order, and the PDB line template - built once per prepared sequence."""
    if "final" not in _state:
        s = sidechain_placeholders[-1].copy()
        order = np.argsort(s.res_nums, kind="mergesort").astype(np.int32)
        s.reorder(order)
        _state["final"] = (s, order, s.pdb_template())
    return _state["final"]


def _final_structure(coords_growth_order):
    from mcsce_b200.structure import _round3_f32

    tmpl, order, _ = _final_template()
    s = tmpl.copy()
    s.xyz = _round3_f32(np.asarray(coords_growth_order, dtype=np.float32))[order]   # 3-decimal table (libstructure.py:157-160)
    return s


class _Conformations(list):
    """``list[Structure | None]`` whose Structure objects are assembled when first looked at: building thousands of
    atom tables in Python would cost more than growing the conformers (the PDB files are already on disk)."""

    def __init__(self, coords, ok):
        super().__init__([None] * len(ok))
        self._coords, self._ok, self._done = coords, ok, [False] * len(ok)

    def _fill(self, i):
        if not self._done[i]:
            self._done[i] = True
            if self._ok[i]:
                super().__setitem__(i, _final_structure(self._coords[i]))
        return super().__getitem__(i)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._fill(k) for k in range(*i.indices(len(self)))]
        return self._fill(i if i >= 0 else i + len(self))

    def __iter__(self):
        return (self._fill(i) for i in range(len(self)))


def _write_pdbs(coords, ok, save_path):
    """All successful conformers of a call to ``<save_path>/<n>.pdb`` through the library's batched writer."""
    import ctypes as C

    from mcsce_b200 import _lib

    tmpl, order, (text, line_len, off) = _final_template()
    lib = _lib.load_library()
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    okb = np.ascontiguousarray(ok, dtype=np.uint8)
    n = C.c_int32(0)
    rc = lib.mcsce_write_pdbs(len(coords), coords.shape[1], coords.ctypes.data_as(C.POINTER(C.c_float)),
                              okb.ctypes.data_as(C.c_void_p), order.ctypes.data_as(C.POINTER(C.c_int32)), text, line_len, off,
                              str(save_path).encode(), 0, C.byref(n))
    if rc != 0:
        raise OSError("could not write every PDB file under %s" % save_path)
    return int(n.value)


def create_side_chain_structure(inputs):
    """``[backbone_coords, beta, retain_idxs, save_addr]`` -> ``(Structure | None, ok, energy | None, save_addr | None)``."""
    backbone_coords, beta, retain_idxs, save_addr = inputs
    coords, energy, ok = _grow(backbone_coords, beta, 1)
    if coords is None:                      # non-root rank under torchrun
        return None, False, None, None
    if not ok[0]:
        return None, False, None, None
    s = _final_structure(coords[0])
    if save_addr is not None:
        s.write_PDB(save_addr)
    return s, True, float(energy[0]), save_addr


def create_side_chain(structure, n_trials, temperature, retain_resi=[], parallel_worker=16, return_first_valid=False):
    """Lowest-energy (or first valid) conformer out of ``n_trials`` grown in one batch."""
    beta = 1 / (temperature * 0.008314462)
    n = int(n_trials)
    if sharding.world()[1] == 1 and n <= MAX_CONFORMERS_PER_LAUNCH and _state.get("hpmf") is None:
        # one rank, one launch: the reduction runs on the device and only the picked conformer is copied to the host
        assert len(sidechain_placeholders) > 0, "Energy functions have not yet initialized!"
        eng, sysm = _state["engine"], _state["system"]
        backbone = np.asarray(structure.coords, dtype=np.float32)
        assert sysm.n_input == backbone.shape[0], \
            "Input structures contain extra sidechain atoms or miss sidechain atoms if fix argument is used"
        key = backbone.tobytes()
        if _state.get("backbone") != key:
            eng.set_backbone(backbone)
            _state["backbone"] = key
        first = _state.get("next_conf", 0)
        _state["next_conf"] = first + n
        out = eng.grow(n, beta, seed=_state.get("seed", 0), conf_id0=first)
        idx, picked = eng.pick(out, first_valid=return_first_valid)
        return None if idx < 0 else _final_structure(picked)
    coords, energy, ok = _grow(structure.coords, beta, n)
    if coords is None or not ok.any():
        return None
    idx = int(np.argmax(ok)) if return_first_valid else int(np.where(ok, energy, np.inf).argmin())
    return _final_structure(coords[idx])


def create_side_chain_ensemble(structure, n_conformations, temperature, save_path, retain_resi=[], parallel_worker=16):
    """``n_conformations`` conformers; writes ``<save_path>/<n>.pdb`` and ``energies.csv``; returns
    ``(list[Structure | None], list[bool])``."""
    beta = 1 / (temperature * 0.008314462)
    coords, energy, ok = _grow(structure.coords, beta, int(n_conformations))
    if coords is None:
        return [], []
    _write_pdbs(coords, ok, save_path)
    success = [bool(k) for k in ok]
    energies = {save_path + f"/{n}.pdb": float(energy[n]) for n in range(int(n_conformations)) if ok[n]}
    conformations = _Conformations(coords, ok)
    with open(save_path + "/energies.csv", "w") as f:
        f.write("File name,Energy(kJ/mol)\n")
        for item in energies:
            f.write("%s,%f\n" % (item, energies[item]))
    return conformations, success
