"""ORACLE - test infrastructure only.  CPU restatement of the MCSCE growth hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product path
(``mcsce_b200``) never does and fails loudly without its CUDA library.

What is restated (NumPy, same dtypes and operation order as the reference so the
results agree with the running reference to the last bit or close to it):

* ``set_chis``            <- ``rotate_sidechain`` / ``rotate_tor``   mcsce/libs/libscbuild.py:17-206
* ``rotate_q``            <- ``rotate_coordinates_Q`` + Hamilton product   mcsce/libs/libcalc.py:358-417
* ``angle_between``       <- ``calc_angle``                            mcsce/libs/libcalc.py:180-203
* ``place_template``      <- ``place_sidechain_template``              mcsce/libs/libcalc.py:419-482
* ``pair_distances``      <- ``calc_new_vs_old_dists``                 mcsce/libs/libcalc.py:47-78
* ``EnergyFunction``      <- ``prepare_energy_function`` closure + ``energycalculator_ij``
                             mcsce/libs/libenergy.py:27-233, 409-537, 680-689, 717-724, 786-810
* ``choose``              <- ``choose_random`` with the uniform passed in   mcsce/core/side_chain_builder.py:33-40
* ``grow_conformer``      <- ``create_side_chain_structure``           mcsce/core/side_chain_builder.py:109-205

PARITY PIN: the reference ships no tests or golden vectors (tests/test_1.py is
empty), so this oracle is pinned against outputs of the *running* reference,
generated in the build container by ``tests/golden/make_golden.py`` and committed
under ``tests/golden/*.npz`` (``tests/test_oracle_golden.py``).

The chemistry tables (force-field parameters, templates, topology) and the
rotamer-library lookup come from ``mcsce_b200.chem`` / ``mcsce_b200.core`` (host
data, checked against the reference in the same tests); all arithmetic of the
hot path is restated here independently of the CUDA host code.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import blas as _blas

from mcsce_b200.chem.tables import (CLASH_FACTOR, COULOMB_K, DIELECTRIC_SCALE, INTER_BOND_1,
                                    INTER_BOND_EQ3, INTER_BOND_LE2, LJ14_EXTRA_SCALE, NO_CHI_RESIDUES,
                                    VDW_RADII, dihedral_f32, load_tables)

F32, F64 = np.float32, np.float64


# ---------------------------------------------------------------------------
# quaternion rotation and template placement
# ---------------------------------------------------------------------------

def rotate_q(coords, rot_vec, angle_rad):
    """Rotate ``coords`` (N,3) about unit ``rot_vec`` by ``angle_rad``; float64 out.

    The reference's jitted version promotes float32 inputs to float64 at the first
    product with the float64 sin/cos, and evaluates the two Hamilton products term by
    term, left to right (libcalc.py:359-366, 401-412)."""
    angle_rad = F64(angle_rad)
    rv = np.asarray(rot_vec).astype(F64)
    s = np.sin(angle_rad / 2)
    b2, b3, b4 = s * rv[0], s * rv[1], s * rv[2]
    b1 = np.cos(angle_rad / 2)
    x, y, z = (np.asarray(coords)[:, k].astype(F64) for k in range(3))
    zero = F64(0.0)
    c1 = (b1 * zero) - (b2 * x) - (b3 * y) - (b4 * z)
    c2 = (b1 * x) + (b2 * zero) + (b3 * z) - (b4 * y)
    c3 = (b1 * y) - (b2 * z) + (b3 * zero) + (b4 * x)
    c4 = (b1 * z) + (b2 * y) - (b3 * x) + (b4 * zero)
    nb2, nb3, nb4 = -b2, -b3, -b4
    d2 = (c1 * nb2) + (c2 * b1) + (c3 * nb4) - (c4 * nb3)
    d3 = (c1 * nb3) - (c2 * nb4) + (c3 * b1) + (c4 * nb2)
    d4 = (c1 * nb4) + (c2 * nb3) - (c3 * nb2) + (c4 * b1)
    return np.stack([d2, d3, d4], axis=1)


def _norm32(v):
    """float32 2-norm the way the jitted reference gets it (BLAS snrm2)."""
    return F32(_blas.snrm2(np.ascontiguousarray(v, dtype=F32)))


def _dot32(a, b):
    return F32(_blas.sdot(np.ascontiguousarray(a, dtype=F32), np.ascontiguousarray(b, dtype=F32)))


def _cross32(a, b):
    a = np.asarray(a, dtype=F32); b = np.asarray(b, dtype=F32)
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dtype=F32)


def angle_between(v1, v2):
    """Angle between two float32 vectors; the dot product is float32, the clamp and arccos
    float64 (libcalc.py:189-203: the three branches unify to float64)."""
    v1u = np.asarray(v1, dtype=F32) / _norm32(v1)
    v2u = np.asarray(v2, dtype=F32) / _norm32(v2)
    d = F64(_dot32(v1u, v2u))
    d = min(max(d, F64(-1.0)), F64(1.0))
    return np.arccos(d)


def place_template(bb, tmpl):
    """float64 (M,3): template ``tmpl`` float32 (M,3) placed on backbone ``bb`` float32 (3,3)."""
    bb = np.asarray(bb, dtype=F32)
    tmpl = np.asarray(tmpl, dtype=F32)
    bbtmp = (bb - bb[1]).astype(F32)
    n_ca, n_ca_t = bbtmp[0], tmpl[0]
    ang1 = angle_between(n_ca, n_ca_t)
    rv = _cross32(n_ca_t, n_ca)
    rvu = rv / _norm32(rv)
    rot1 = rotate_q(tmpl, rvu, ang1).astype(F32)
    cross_cnf = _cross32(bbtmp[0], bbtmp[2])
    cross_ss = _cross32(rot1[0], rot1[2])
    ang2 = angle_between(cross_ss, cross_cnf)
    rv = _cross32(cross_ss, cross_cnf)
    rvu = rv / _norm32(rv)
    rot2 = rotate_q(rot1, rvu, ang2)
    return rot2 + bb[1].astype(F64)


def set_chis(template, tors_deg):
    """float32 (n_tmpl,3): residue template with chi_1..chi_n set to ``tors_deg`` degrees.

    Sequential, each rotation applied to the float32 result of the previous one; the
    current dihedral is re-measured in float32 before every rotation; the rotation angle
    ``measured - target`` is wrapped once into (-pi, pi]; rotation about minus the unit
    axis through the pivot atom (libscbuild.py:44-54, 103-206)."""
    tors = np.asarray(tors_deg, dtype=F64) * np.pi / 180
    xyz = template.coords.astype(F32, copy=True)
    for k in range(min(len(tors), len(template.chis))):
        chi = template.chis[k]
        measured = dihedral_f32(xyz[list(chi["dihedral"])])[0]          # float32 scalar
        a0, a1 = chi["axis"]
        if k == 0:
            pivot = xyz[4].copy()
            unit = pivot / np.linalg.norm(pivot)                         # CA is the origin
        else:
            pivot = xyz[a1].copy()
            diff = xyz[a1] - xyz[a0]
            unit = diff / np.linalg.norm(diff)
        idx = chi["moving"]
        displaced = xyz[idx] - pivot                                     # float32
        rot = measured - tors[k]                                         # float64
        if rot > np.pi:
            rot -= 2 * np.pi
        elif rot < -np.pi:
            rot += 2 * np.pi
        xyz[idx] = rotate_q(displaced, -unit, rot) + pivot               # float64 -> stored float32
    return xyz


def trial_side_chain(template, bb, tors_deg):
    """float32 (n_sc,3) coordinates of one trial: chi setting, placement, side-chain rows,
    stored to float32 like ``all_coords`` (side_chain_builder.py:179-181)."""
    placed = place_template(bb, set_chis(template, tors_deg))
    return placed[template.sc_idx].astype(F32)


# ---------------------------------------------------------------------------
# pair distances and the energy function
# ---------------------------------------------------------------------------

def pair_distances(new, old):
    """float64 (T, P) distances: new-new upper triangle (row-major) then new x old (new-major).
    Differences, squares and their sum are float32, the square root float64
    (libcalc.py:64-67, 73-76)."""
    new = np.asarray(new, dtype=F32); old = np.asarray(old, dtype=F32)
    n = new.shape[1]
    iu, ju = np.triu_indices(n, k=1)
    d_nn = new[:, ju, :] - new[:, iu, :]
    d_no = old[:, None, :, :] - new[:, :, None, :]
    d_no = d_no.reshape(new.shape[0], -1, 3)
    d = np.concatenate([d_nn, d_no], axis=1)
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    sq = x * x + y * y + z * z                                              # float32
    return sq.astype(F64) ** 0.5


def _relation(tables, a_name, a_num, b_name, b_num, label_of_num):
    """0 none / 1 bonded / 2 two bonds / 3 three bonds, per the reference's rule-based masks
    (libenergy.py:286-345; inter-residue rules build_definitions.py:460-482)."""
    if a_num == b_num:
        return tables.bond_distance(label_of_num[a_num], a_name, b_name)
    if b_num == a_num + 1:
        lo, hi = a_name, b_name
    elif b_num == a_num - 1:
        lo, hi = b_name, a_name
    else:
        return 0
    if hi in INTER_BOND_1.get(lo, ()):
        return 1
    if hi in INTER_BOND_LE2.get(lo, ()):
        return 2
    if hi in INTER_BOND_EQ3.get(lo, ()):
        return 3
    return 0


class EnergyFunction:
    """``calculate(coords_new, coords_old) -> (T,) float64`` with ``inf`` for clashing trials."""

    def __init__(self, atom_labels, res_nums, res_labels, n_term, c_term, new_idx=None, old_idx=None,
                 terms=("lj", "clash", "coulomb"), tables=None, batch_size=16):
        t = tables or load_tables()
        atom_labels = np.asarray(atom_labels); res_nums = np.asarray(res_nums).astype(int)
        res_labels = np.asarray(res_labels)
        n_all = len(atom_labels)
        if new_idx is None:
            new_idx, old_idx = np.arange(n_all), np.array([], dtype=int)
        new_idx = np.asarray(new_idx, dtype=int); old_idx = np.asarray(old_idx, dtype=int)
        self.terms = tuple(terms)
        self.batch_size = batch_size
        label_of_num = {}
        for num, lab in zip(res_nums, res_labels):
            label_of_num.setdefault(int(num), str(lab))
        q = np.empty(n_all); s = np.empty(n_all); e = np.empty(n_all); r = np.empty(n_all)
        for a in range(n_all):
            q[a], s[a], e[a] = t.atom_params(str(res_labels[a]), str(atom_labels[a]), bool(n_term[a]), bool(c_term[a]))
            r[a] = VDW_RADII[str(atom_labels[a])[0]]
        nn = len(new_idx)
        iu, ju = np.triu_indices(nn, k=1)
        pi = np.concatenate([new_idx[iu], np.repeat(new_idx, len(old_idx))])
        pj = np.concatenate([new_idx[ju], np.tile(old_idx, nn)])
        # per-pair parameters with the reference's expressions (libenergy.py:462-481, 518-522, 526-537)
        eps_ij = (e[pi] * e[pj]) ** 0.5
        sig_ij = (s[pi] + s[pj]) * 5
        self.acoeff = 4 * eps_ij * (sig_ij ** 12)
        self.bcoeff = 4 * eps_ij * (sig_ij ** 6)
        qq = q[pi] * q[pj]
        qq *= DIELECTRIC_SCALE
        qq *= COULOMB_K
        self.charges = qq
        self.thr = (r[pi] + r[pj]) * CLASH_FACTOR
        # bonded relations: only pairs in the same or adjacent residue numbers can be related
        near = np.where(np.abs(res_nums[pi] - res_nums[pj]) <= 1)[0]
        rel = np.zeros(len(pi), dtype=np.int8)
        for p in near:
            a, b = pi[p], pj[p]
            lo, hi = (a, b) if a < b else (b, a)       # masks are keyed (smaller index, larger index)
            rel[p] = _relation(t, str(atom_labels[lo]), int(res_nums[lo]), str(atom_labels[hi]),
                               int(res_nums[hi]), label_of_num)
        eq3, le2, b1 = rel == 3, (rel == 1) | (rel == 2), rel == 1
        self.acoeff[eq3] *= t.lj14scale * LJ14_EXTRA_SCALE
        self.bcoeff[eq3] *= t.lj14scale * LJ14_EXTRA_SCALE
        self.acoeff[le2] = np.nan
        self.bcoeff[le2] = np.nan
        self.charges[eq3] *= t.coulomb14scale
        self.charges[le2] = np.nan
        self.thr[b1] = 0
        self.n_pairs = len(pi)

    def calculate(self, coords_new, coords_old):
        assert coords_new.ndim == 3
        T = len(coords_new)
        out = np.zeros(T)
        for s0 in range(0, T, self.batch_size):
            d = pair_distances(coords_new[s0:s0 + self.batch_size], coords_old[s0:s0 + self.batch_size])
            if "clash" in self.terms:
                clash = (d < self.thr[None]).any(axis=1)
            else:
                clash = np.zeros(len(d), dtype=bool)
            e = np.zeros(len(d))
            ok = d[~clash]
            with np.errstate(divide="ignore", invalid="ignore"):
                if "lj" in self.terms:
                    e[~clash] += np.nansum(self.acoeff[None] / ok ** 12 - self.bcoeff[None] / ok ** 6, axis=1)
                if "coulomb" in self.terms:
                    e[~clash] += np.nansum(self.charges[None] / ok, axis=1)
            e[clash] = np.inf
            out[s0:s0 + len(d)] = e
        return out

    __call__ = calculate


# ---------------------------------------------------------------------------
# selection and the growth loop
# ---------------------------------------------------------------------------

def choose(weights, u):
    """Index of the first running sum exceeding ``u * total``; sequential float64 sums;
    last index if none does (side_chain_builder.py:33-40, uniform supplied by the caller)."""
    cum = np.empty(len(weights))
    acc = 0.0
    for i, w in enumerate(weights):
        acc = acc + float(w)
        cum[i] = acc
    pointer = float(u) * cum[-1]
    for i, c in enumerate(cum):
        if c > pointer:
            return i
    return len(weights) - 1


def prepare_energy_functions(structure, retain_idxs=(), terms=("lj", "clash", "coulomb"), tables=None):
    """{level: EnergyFunction}: level 0 = input atoms alone, level i+1 = residue i's side chain
    against everything placed before it (initialize_func_calc, side_chain_builder.py:42-107)."""
    t = tables or load_tables()
    work = structure.copy()
    n_term, c_term = work.get_terminal_res_atom_arr()
    out = {0: EnergyFunction(work.atom_labels, work.res_nums, work.res_labels, n_term, c_term, terms=terms, tables=t)}
    first = int(work.res_nums[0])
    for i, (res, chain) in enumerate(zip(work.residue_types, work.residue_chains)):
        if first + i in retain_idxs:
            continue
        tmpl = t.templates[res]
        work.append_atoms([tmpl.names[k] for k in tmpl.sc_idx], tmpl.resname, chain, first + i,
                          np.zeros((tmpl.n_sc, 3), dtype=F32))
        if res in NO_CHI_RESIDUES:
            continue
        n_all = len(work)
        n_term, c_term = work.get_terminal_res_atom_arr()
        out[i + 1] = EnergyFunction(work.atom_labels, work.res_nums, work.res_labels, n_term, c_term,
                                    np.arange(n_all - tmpl.n_sc, n_all), np.arange(n_all - tmpl.n_sc), terms=terms, tables=t)
    return out


def grow_conformer(structure, backbone_coords, beta, chi_for_step, uniform_for_step, retain_idxs=(),
                   terms=("lj", "clash", "coulomb"), tables=None, efuncs=None, record=None, extra_term=None):
    """One conformer of the growth loop (side_chain_builder.py:109-205).

    ``structure``: backbone-only ``mcsce_b200.structure.Structure`` (atom bookkeeping only).
    ``chi_for_step(i, resname, phi, psi) -> (chis (T,nchi) deg, probs (T,))``: the *perturbed*
    trial set of residue ordinal i (library rows + noise, supplied by the caller so that the
    oracle, the reference and the GPU path see identical trials).
    ``uniform_for_step(i) -> u`` in [0,1).
    ``extra_term(atom_labels, res_nums, res_labels, coords) -> float``: a coordinate-based term on the
    level's whole structure (the reference's ``efuncs_coords``, libenergy.py:806-808), added to every trial
    that passed the clash filter and to the starting energy.
    Returns dict(ok, energy, coords (growth order, float32), records).
    """
    from mcsce_b200.structure import _round3_f32
    from mcsce_b200.system import backbone_torsions_deg

    t = tables or load_tables()
    backbone_coords = np.asarray(backbone_coords, dtype=F32)
    work = structure.copy()
    assert len(work) == backbone_coords.shape[0]
    work.xyz = _round3_f32(backbone_coords)
    ncac = work.get_sorted_minimal_backbone_coords()
    phi, psi = backbone_torsions_deg(ncac)
    res_types = work.residue_types
    res_chains = work.residue_chains
    first = int(work.res_nums[0])
    cache = efuncs if efuncs is not None else {}

    def efunc(level, new_idx=None, old_idx=None):
        if level not in cache:
            n_term, c_term = work.get_terminal_res_atom_arr()
            cache[level] = EnergyFunction(work.atom_labels, work.res_nums, work.res_labels, n_term, c_term,
                                          new_idx, old_idx, terms=terms, tables=t)
        return cache[level]

    coords = backbone_coords.copy()
    energy = efunc(0).calculate(coords[None], coords[None, :0])[0]
    if extra_term is not None:
        energy = energy + extra_term(work.atom_labels, work.res_nums, work.res_labels, coords)
    for i, res in enumerate(res_types):
        num = first + i
        if num in retain_idxs:
            continue
        tmpl = t.templates[res]
        n_sc = tmpl.n_sc
        bb = ncac[3 * i: 3 * i + 3]
        # atom bookkeeping of add_side_chain: block appended at the end
        if len(work) == len(coords):
            work.append_atoms([tmpl.names[k] for k in tmpl.sc_idx], tmpl.resname, res_chains[i], num,
                              np.zeros((n_sc, 3), dtype=F32))
        if res in NO_CHI_RESIDUES:
            placed = place_template(bb, tmpl.coords)[tmpl.sc_idx].astype(F32)
            coords = np.concatenate([coords, placed])
            continue
        n_all = len(coords) + n_sc
        ef = efunc(i + 1, np.arange(n_all - n_sc, n_all), np.arange(n_all - n_sc))
        chis, probs = chi_for_step(i, res, phi[i], psi[i])
        T = len(chis)
        trial = np.stack([trial_side_chain(tmpl, bb, chis[k]) for k in range(T)])
        env = np.broadcast_to(coords[None], (T,) + coords.shape)
        e = ef.calculate(trial, env)
        if extra_term is not None:
            for k in range(T):
                if np.isfinite(e[k]):
                    e[k] += extra_term(work.atom_labels, work.res_nums, work.res_labels,
                                       np.concatenate([coords, trial[k]]))
        if record is not None:
            record.append(dict(step=i, res=res, chis=np.array(chis, dtype=F64), probs=np.array(probs),
                               trial=trial.copy(), energies=e.copy(), n_env=len(coords)))
        if np.isinf(e).all():
            return dict(ok=False, energy=None, coords=None, failed_at=i)
        emin = e.min()
        w = np.asarray(probs) * np.exp(-beta * (e - emin))
        u = uniform_for_step(i)
        sel = choose(w, u)
        if record is not None:
            record[-1].update(u=u, selected=sel, weights=w)
        coords = np.concatenate([coords, trial[sel]])
        energy = energy + ((e[sel] - emin) + emin)
    return dict(ok=True, energy=float(energy), coords=coords)
