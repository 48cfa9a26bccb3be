"""Drop-in for ``mcsce.libs.libenergy.prepare_energy_function`` (libenergy.py:27-233).

Same signature and the same returned closure contract - ``calculate(coords_new (T, n_new, 3),
coords_old (T, n_old, 3)) -> (T,) float64`` with ``inf`` where the clash filter fires
(libenergy.py:786-810) - but the pair sums run in the CUDA scoring kernels (K2/K2s) through
``libmcsce_b200.so``.  No O(N^2) masks are built: the labels are turned into per-atom classes and
the exact special-pair list of the new side chain (mcsce_b200.system).

Terms: ``lj``, ``clash``, ``coulomb`` - the set the reference CLI uses (cli.py:149-151) - and ``hpmf``.
The reference's angle / dihedral / gb / hpmf switches cannot run in the reference itself
(``connectivity_matrix`` is commented out at libenergy.py:110-113, keras is absent); asking for
angle / dihedral / gb here raises ``NotImplementedError`` instead of ``UnboundLocalError``.

``hpmf`` is the one of them whose building blocks run in the reference when called directly
(libhpmf.py, libsasa.py).  It is wired the way the reference's code reads once its two slips are
repaired: the bond connectivity is the 1-bond mask (the commented-out lines), and the closure gets the
level's whole structure, new and old atoms in label order (``efuncs_coords``, libenergy.py:806-808):
every trial that passes the clash filter gets ``V_hpmf`` of (old atoms + its new atoms) added.
"""
from __future__ import annotations

import numpy as np

_SUPPORTED = ("lj", "clash", "coulomb", "hpmf")


def _device():
    import os
    return int(os.environ.get("LOCAL_RANK", os.environ.get("MCSCE_B200_DEVICE", "0")))


def prepare_energy_function(atom_labels, residue_numbers, residue_labels, N_terminal_indicators, C_terminal_indicators,
                            forcefield=None, batch_size=16, partial_indices=None, terms=None, angle_term=True,
                            dihedral_term=True, clash_term=True, lj_term=True, coulomb_term=True, gb_term=True,
                            hpmf_term=True):
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.system import build_system_from_labels

    if terms is None:
        flags = dict(angle=angle_term, dihedral=dihedral_term, clash=clash_term, lj=lj_term, coulomb=coulomb_term,
                     gb=gb_term, hpmf=hpmf_term)
        terms = [k for k, v in flags.items() if v]
    bad = [t for t in terms if t not in _SUPPORTED]
    if bad:
        raise NotImplementedError("energy terms %s are not runnable in the reference either (libenergy.py:110-113, "
                                  "libgb.py:14); supported: %s" % (bad, list(_SUPPORTED)))
    sysm = build_system_from_labels(atom_labels, residue_numbers, residue_labels, N_terminal_indicators,
                                    C_terminal_indicators, partial_indices, terms=tuple(terms))
    eng = GrowthEngine(_device()).set_system(sysm)
    eng.set_backbone(np.zeros((sysm.n_input, 3), dtype=np.float32))
    hpmf = None
    if "hpmf" in terms:
        from mcsce_b200.hpmf import calculator_for_labels
        hpmf = calculator_for_labels(atom_labels, residue_numbers, residue_labels, device=_device())

    def add_hpmf(out, coords_new, coords_old):
        ok = np.isfinite(out)
        if hpmf is None or not ok.any():
            return out
        if partial_indices is None:
            full = coords_new[ok]
        else:
            new_idx, old_idx = (np.asarray(x, dtype=np.int64) for x in partial_indices)
            full = np.empty((int(ok.sum()), len(atom_labels), 3), dtype=np.float32)
            full[:, new_idx] = coords_new[ok]
            full[:, old_idx] = coords_old[ok]
        out[ok] += hpmf.energy(full).cpu().numpy()
        return out

    def calculate(coords_new, coords_old):
        coords_new = np.asarray(coords_new, dtype=np.float32)
        coords_old = np.asarray(coords_old, dtype=np.float32)
        return add_hpmf(_pair_terms(coords_new, coords_old), coords_new, coords_old)

    def _pair_terms(coords_new, coords_old):
        coords_new = np.asarray(coords_new, dtype=np.float32)
        coords_old = np.asarray(coords_old, dtype=np.float32)
        assert len(coords_new.shape) == 3
        T = len(coords_new)
        if T == 0:
            return np.zeros(0)
        if partial_indices is None:
            out = np.empty(T)
            for t in range(T):                      # whole-structure energy: every atom is "new"
                eng.set_backbone(coords_new[t])
                out[t] = eng.input_energy()
            return out
        if T > 1 and (coords_old == coords_old[:1]).all():
            return eng.score_trials(0, coords_new[None], coords_old[:1]).cpu().numpy()[0]
        return eng.score_trials(0, coords_new[:, None], coords_old).cpu().numpy()[:, 0]

    calculate.engine = eng
    return calculate
