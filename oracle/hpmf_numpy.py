"""ORACLE (test infrastructure, not product code) - NumPy restatement of the reference's
hydrophobic-PMF term with the reference's dtypes and operation order.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file.

Follows (paths relative to the reference's ``src/mcsce/``):

* ``libs/libsasa.py:76-109``  Hasel-Hendrickson-Still approximate SASA, product formula;
* ``libs/libhpmf.py:22-68``   atom labels -> SASA atom types (``convert_atom_types``);
* ``libs/libhpmf.py:70-111``  three-Gaussian PMF over carbon pairs with SA > 6 A^2, tanh(SA) weights;
* ``libs/libenergy.py:97-113`` (commented-out intent) bond connectivity = the 1-bond mask.

**Parity pin**: the reference cannot run the term through ``prepare_energy_function`` (SURVEY TL;DR 3),
and its test suite holds nothing for it; the pin is the output of the reference's own ``calc_sasa`` /
``init_hpmf_calculator`` called directly on the reference-grown structures of the growth fixtures
(``tests/golden/hpmf.npz``, made by ``tests/golden/make_golden_hpmf.py``), checked in
``tests/test_oracle_hpmf.py``.  How the term would enter the per-trial energy is *not* pinned: the
reference's wiring (``libenergy.py:213-220, 807``) raises before it gets there.

Memory: the reference materialises several N x N float64 arrays; this restatement walks row blocks,
same arithmetic per element.
"""
import numpy as np


def convert_atom_types(atom_labels, residue_numbers, residue_labels, type_map):
    """libhpmf.py:22-68.  ``type_map``: {residue: {atom: type or None}} (core/definitions.py:181-511)."""
    keep, types = [], []
    for label, resnum, resname in zip(atom_labels, residue_numbers, residue_labels):
        t = None
        if resnum == 1 and label == "N":
            t = "N-SP3"
        elif resnum == 1 and label in ("H1", "H2", "H3"):
            t = "H-NH+"
        elif label == "OXT":
            t = "O-"
        else:
            t = type_map[resname][label]          # KeyError for residues the reference has no types for
        keep.append(t is not None)
        if t is not None:
            types.append(t)
    return np.array(keep), types


def sasa(coords, types, bonded, P, R, pij_bonded, pij_non_bonded, r_solvent=1.4, block=256):
    """libsasa.py:76-109.  ``coords`` (N, 3) float32 of the typed atoms, ``bonded`` (N, N) bool."""
    x = np.asarray(coords)
    N = len(types)
    Ri = np.array([R[t] for t in types])
    Pi = np.array([P[t] for t in types])
    Si = 4 * np.pi * (Ri + r_solvent) ** 2
    out = np.empty(N)
    for lo in range(0, N, block):
        hi = min(N, lo + block)
        # float32 differences and norm (np.linalg.norm on float32 stays float32)
        diff = x[lo:hi, None, :] - x[None, :, :]
        dij = np.sqrt(np.add.reduce(diff * diff, axis=-1))
        thr = Ri[lo:hi, None] + Ri[None, :] + 2 * r_solvent
        mask = dij < thr
        mask[np.arange(hi - lo), np.arange(lo, hi)] = False
        Pij = np.where(bonded[lo:hi], pij_bonded, pij_non_bonded)
        bij = np.pi * (Ri + r_solvent)[None, :] * (thr - dij) * (1 + (Ri[None, :] - Ri[lo:hi, None]) / (dij + 1e-8))
        terms = 1 - Pi[lo:hi, None] * Pij * bij / Si[lo:hi, None]
        terms[~mask] = 1
        out[lo:hi] = Si[lo:hi] * np.prod(terms, axis=-1)
    return out


def hpmf(coords, atom_filter, types, bonded, tables, threshold=None):
    """libhpmf.py:92-109 on one structure: ``coords`` (N_all, 3) float32."""
    C, W, H = tables["hpmf"]["C"], tables["hpmf"]["W"], tables["hpmf"]["H"]
    threshold = tables["hpmf"]["AC"] if threshold is None else threshold
    x = np.asarray(coords)[atom_filter]
    sa = sasa(x, types, bonded, tables["P"], tables["R"], tables["pij_bonded"], tables["pij_non_bonded"],
              tables["r_solvent"])
    carbon = np.array([t.startswith("C") for t in types])
    csa = sa[carbon]
    cx = x[carbon]
    sel = csa > threshold
    cx = cx[sel]
    diff = cx[:, None, :] - cx[None, :, :]
    rij = np.sqrt(np.add.reduce(diff * diff, axis=-1))
    g = (H[0] * np.exp(-((rij - C[0]) / W[0]) ** 2) + H[1] * np.exp(-((rij - C[1]) / W[1]) ** 2)
         + H[2] * np.exp(-((rij - C[2]) / W[2]) ** 2))
    pre = np.tanh(csa[sel, None]) * np.tanh(csa[None, sel])
    return 0.5 * np.sum(g * pre * (1 - np.eye(int(sel.sum())))), sa
