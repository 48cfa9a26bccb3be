"""Host-side tables for one backbone: everything the CUDA growth path needs, in O(N).

Replaces the reference's ``initialize_func_calc`` setup
(``mcsce/core/side_chain_builder.py:42-107``), which builds L+1 placeholder
structures and L+1 energy closures, each closure re-deriving three N(N-1)/2
bonded masks (``mcsce/libs/libenergy.py:48-106``).  The same information is held
here as per-atom parameters plus, per growth step, a short list of *special*
pairs (everything within three bonds of the new side chain, and the side
chain's own internal pairs) with their exact per-pair coefficients; every other
pair is generic and is scored from per-atom values.

Atom order ("growth order") is the reference's: the input atoms in file order,
then one side-chain block per residue in residue order
(``libstructure.py:498-514`` appends each block at the end).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from mcsce_b200.chem.tables import (CLASH_ELEMENTS, CLASH_FACTOR, COULOMB_K, DIELECTRIC_SCALE,
                                    INTER_BOND_1, INTER_BOND_EQ3, INTER_BOND_LE2, LJ14_EXTRA_SCALE,
                                    NO_CHI_RESIDUES, VDW_RADII, load_tables)
from mcsce_b200.structure import _round3_f32

REL_NONE, REL_BOND1, REL_LE2, REL_EQ3 = 0, 1, 2, 3
MAX_CHI = 6          # widest trial row in the libraries (KCX carries six, five are applied)
MAX_TMPL = 32        # template atoms (M3L has 31): one warp lane per template atom in the placement kernel


def backbone_torsions_deg(ncac):
    """phi/psi in degrees per residue from the interleaved N,CA,C float32 array, first phi and
    last psi NaN - the quantities side_chain_builder.py:144-147 derives through
    libcalc.py:206-304 (float32 vectors, ``-arctan2(sin, cos)``)."""
    crds = np.asarray(ncac, dtype=np.float32)
    q = crds[1:] - crds[:-1]
    cross = np.cross(q[:-1], q[1:])
    unitary = cross / np.linalg.norm(cross, axis=1, keepdims=True)
    u0, u1 = unitary[:-1], unitary[1:]
    u3 = q[1:-1] / np.linalg.norm(q[1:-1], axis=1, keepdims=True)
    u2 = np.cross(u3, u1)
    cos_t = (u0 * u1).sum(1, dtype=np.float32)
    sin_t = (u0 * u2).sum(1, dtype=np.float32)
    dih = -np.arctan2(sin_t, cos_t)
    phi = np.concatenate([[np.nan], dih[2::3]]) * 180 / np.pi
    psi = np.concatenate([dih[::3], [np.nan]]) * 180 / np.pi
    return phi, psi


@dataclass
class GrowthStep:
    index: int                  # residue ordinal
    resnum: int
    resname: str                # backbone label == template key
    kind: str                   # 'grow' | 'rigid' (GLY/ALA) | 'skip' (retained)
    sc_start: int = 0           # first growth-order index of this residue's side-chain block
    n_sc: int = 0
    bb_rows: tuple = ()         # growth-order indices of this residue's N, CA, C
    phi: float = np.nan
    psi: float = np.nan
    n_chi_lib: int = 0          # chi columns in the trial rows
    n_chi: int = 0              # chi rotations actually applied
    chi_mean: np.ndarray = None  # (T, n_chi_lib) degrees
    chi_sigma: np.ndarray = None
    prob: np.ndarray = None      # (T,)
    special_env: np.ndarray = None   # growth-order indices of environment atoms scored pair-by-pair
    # special pair table: rows (new k, partner, A, B, Q, thr, flags); partner >= 0 -> index into
    # special_env, partner < 0 -> -(k'+1): another new atom of the same side chain
    sp_k: np.ndarray = None
    sp_partner: np.ndarray = None
    sp_coef: np.ndarray = None   # (n_pairs, 4) float64: A, B, Q, thr ; excluded pairs carry A=B=Q=0


@dataclass
class GrowthSystem:
    names: np.ndarray
    resnames: np.ndarray
    resnums: np.ndarray
    chains: np.ndarray
    res_index: np.ndarray
    is_nterm: np.ndarray
    is_cterm: np.ndarray
    charge: np.ndarray
    sigma: np.ndarray
    eps: np.ndarray
    radius: np.ndarray
    appear: np.ndarray           # -1 for input atoms, residue ordinal for grown atoms
    n_input: int
    steps: list = field(default_factory=list)
    terms: tuple = ("lj", "clash", "coulomb")
    # generic-pair classes and the slot layout of the device environment
    cls_of_atom: np.ndarray = None
    n_cls: int = 0
    cls_s2: np.ndarray = None    # (n_cls, n_cls) float64 ((s_i+s_j)*5)^2
    cls_e4: np.ndarray = None    # (n_cls, n_cls) float64 4*sqrt(e_i e_j)
    cls_thr: np.ndarray = None   # (n_cls, n_cls) float64 clash distance
    slot_of_atom: np.ndarray = None
    atom_of_slot: np.ndarray = None
    cls_start: np.ndarray = None   # (n_cls+1,) slot offsets
    cls_active: np.ndarray = None  # (L+1, n_cls) active atoms of each class when residue i is grown
    bb_pairs: tuple = None       # backbone self-energy special pairs (i, j, coef)
    rotamer_library: object = None   # libraries the trial sets come from (None: scoring-only system)
    ptm_library: object = None
    trial_key: bytes = None          # N/CA/C coordinates the current trial sets were made for

    @property
    def n_atoms(self):
        return len(self.names)

    @property
    def n_res(self):
        return len(self.steps)


def relation(tables, a_name, a_resnum, a_label, b_name, b_resnum, b_label):
    """Bonded relation of an atom pair under the reference's rule-based masks (Q4):
    intra-residue by template topology of the residue's *backbone* label, inter-residue only
    n -> n+1 by fixed atom-name rules (libenergy.py:286-345, build_definitions.py:460-482)."""
    if a_resnum == b_resnum:
        d = tables.bond_distance(a_label, a_name, b_name)
        return (REL_NONE, REL_BOND1, REL_LE2, REL_EQ3)[d]
    if b_resnum == a_resnum - 1:
        a_name, b_name = b_name, a_name
    elif b_resnum != a_resnum + 1:
        return REL_NONE
    # from here a_name is in residue n and b_name in residue n+1
    if b_name in INTER_BOND_1.get(a_name, ()):
        return REL_BOND1
    if b_name in INTER_BOND_LE2.get(a_name, ()):
        return REL_LE2
    if b_name in INTER_BOND_EQ3.get(a_name, ()):
        return REL_EQ3
    return REL_NONE


def pair_coefficients(tables, rel, qi, si, ei, ri, qj, sj, ej, rj, terms):
    """(A, B, Q, thr) of one pair, float64, evaluated with the reference's expressions
    (libenergy.py:474-481 LJ mixing, :519-522 Coulomb, :151-153 clash, :171-178/199-200 scaling)."""
    eps_ij = (ei * ej) ** 0.5
    sig_ij = (si + sj) * 5
    A = 4 * eps_ij * (sig_ij ** 12)
    B = 4 * eps_ij * (sig_ij ** 6)
    Q = qi * qj
    Q *= DIELECTRIC_SCALE
    Q *= COULOMB_K
    thr = (ri + rj) * CLASH_FACTOR
    if rel == REL_EQ3:
        A *= tables.lj14scale * LJ14_EXTRA_SCALE
        B *= tables.lj14scale * LJ14_EXTRA_SCALE
        Q *= tables.coulomb14scale
    if rel in (REL_BOND1, REL_LE2):
        A = B = Q = 0.0
    if rel == REL_BOND1:
        thr = 0.0
    if "lj" not in terms:
        A = B = 0.0
    if "coulomb" not in terms:
        Q = 0.0
    if "clash" not in terms:
        thr = 0.0
    return A, B, Q, thr


def build_system(structure, retain_idxs=(), terms=("lj", "clash", "coulomb"), rotamer_library=None,
                 ptm_library=None, tables=None):
    """GrowthSystem for a backbone-only (plus retained residues) structure."""
    tables = tables or load_tables()
    terms = tuple(terms)
    names = [str(n) for n in structure.cols["name"]]
    resnames = [str(n) for n in structure.cols["resname"]]
    resnums = [int(n) for n in structure.cols["resseq"]]
    chains = [str(c) for c in structure.cols["chain"]]
    n_input = len(names)
    res_types = structure.residue_types
    res_chains = structure.residue_chains
    first_num = resnums[0]
    ordinal_of_num = {}
    # residue ordinal = position in residue_types; the reference maps ordinal <-> number as
    # ordinal + res_nums[0] (side_chain_builder.py:88-90,153) i.e. contiguous numbering
    for i in range(len(res_types)):
        ordinal_of_num[first_num + i] = i
    res_index = [ordinal_of_num.get(n, -1) for n in resnums]
    if min(res_index) < 0:
        raise ValueError("residue numbers must be contiguous and unique across chains")
    # every (chain, number) of the input must be one residue of the ordered list: chains that restart their numbering
    # (A:1-10, B:1-10) would pass the test above and silently alias two residues onto one ordinal
    seen = []
    for key in zip(chains, resnums):
        if not seen or seen[-1] != key:
            seen.append(key)
    if len(set(seen)) != len(res_types) or [n for _, n in sorted(set(seen), key=lambda k: k[1])] != list(
            range(first_num, first_num + len(res_types))):
        raise ValueError("residue numbers must be contiguous and unique across chains (found %d distinct (chain, number) "
                         "pairs for %d residues; renumber chains consecutively)" % (len(set(seen)), len(res_types)))
    appear = [-1] * n_input

    ncac = structure.get_sorted_minimal_backbone_coords()
    phi, psi = backbone_torsions_deg(ncac)
    name_arr = np.array(names)
    n_rows, ca_rows, c_rows = (np.where(name_arr == a)[0] for a in ("N", "CA", "C"))
    if not (len(n_rows) == len(ca_rows) == len(c_rows) == len(res_types)):
        raise ValueError("every residue needs exactly one N, CA and C atom (found %d/%d/%d for %d residues)"
                         % (len(n_rows), len(ca_rows), len(c_rows), len(res_types)))

    steps = []
    for i, (res, chain) in enumerate(zip(res_types, res_chains)):
        num = first_num + i
        st = GrowthStep(index=i, resnum=num, resname=res, kind="grow",
                        bb_rows=(int(n_rows[i]), int(ca_rows[i]), int(c_rows[i])),
                        phi=float(phi[i]), psi=float(psi[i]))
        if num in retain_idxs:
            st.kind = "skip"
            steps.append(st)
            continue
        tmpl = tables.templates[res]
        st.sc_start, st.n_sc = len(names), tmpl.n_sc
        for k in tmpl.sc_idx:
            names.append(tmpl.names[k]); resnames.append(tmpl.resname); resnums.append(num)
            chains.append(chain); res_index.append(i); appear.append(i)
        if res in NO_CHI_RESIDUES:
            st.kind = "rigid"
        steps.append(st)

    n = len(names)
    sysm = GrowthSystem(
        names=np.array(names), resnames=np.array(resnames), resnums=np.array(resnums, dtype=np.int64),
        chains=np.array(chains), res_index=np.array(res_index, dtype=np.int64),
        is_nterm=np.zeros(n, dtype=bool), is_cterm=np.zeros(n, dtype=bool),
        charge=np.zeros(n), sigma=np.zeros(n), eps=np.zeros(n), radius=np.zeros(n),
        appear=np.array(appear, dtype=np.int64), n_input=n_input, steps=steps, terms=terms)

    # terminal flags: first / last residue of every chain, keyed by (chain, number) (libstructure.py:163-183)
    chain_res = {}
    for ch, num in zip(chains, resnums):
        chain_res.setdefault(ch, {}).setdefault(num, None)
    for ch, d in chain_res.items():
        keys = list(d.keys())
        ch_mask = sysm.chains == ch
        sysm.is_nterm |= ch_mask & (sysm.resnums == keys[0])
        sysm.is_cterm |= ch_mask & (sysm.resnums == keys[-1])
    sysm.rotamer_library, sysm.ptm_library = rotamer_library, ptm_library
    if rotamer_library is not None:
        assign_trials(sysm, structure.xyz[:n_input], tables)
    _finalize(sysm, tables)
    return sysm


def assign_trials(sysm, input_coords, tables=None):
    """phi/psi and the trial set of every grown residue for these input-atom coordinates.  The reference derives both
    from the coordinates handed to every ``create_side_chain_structure`` call (side_chain_builder.py:143-147, 171), so
    one prepared sequence serves any backbone conformation; ``GrowthEngine.set_backbone`` calls this when the backbone
    torsion atoms differ from the ones the current tables were made for.  Returns True if anything was recomputed."""
    tables = tables or load_tables()
    rounded = _round3_f32(np.asarray(input_coords, dtype=np.float32))     # read back from the 3-decimal table (:143-144)
    ncac = np.ascontiguousarray(rounded[_bb_rows(sysm)].reshape(-1, 3))
    key = ncac.tobytes()
    if sysm.trial_key == key:
        return False
    phi, psi = backbone_torsions_deg(ncac)
    for st in sysm.steps:
        st.phi, st.psi = float(phi[st.index]), float(psi[st.index])
        if st.kind != "grow":
            continue
        mean, sig, prob = sysm.rotamer_library.trial_distribution(st.resname, st.phi, st.psi, sysm.ptm_library)
        st.chi_mean, st.chi_sigma, st.prob = np.asarray(mean, float), np.asarray(sig, float), np.asarray(prob, float)
        st.n_chi_lib = st.chi_mean.shape[1]
        st.n_chi = min(st.n_chi_lib, tables.templates[st.resname].n_sound)
    sysm.trial_key = key
    return True


def _finalize(sysm, tables):
    """Per-atom parameters, special-pair lists, atom classes and slot layout."""
    n = sysm.n_atoms
    names, resnames, resnums = sysm.names, sysm.resnames, sysm.resnums
    for a in range(n):
        q, s, e = tables.atom_params(str(resnames[a]), str(names[a]), bool(sysm.is_nterm[a]), bool(sysm.is_cterm[a]))
        sysm.charge[a], sysm.sigma[a], sysm.eps[a] = q, s, e
        sysm.radius[a] = VDW_RADII[str(names[a])[0]]
    # backbone label per residue number (bonds_intra is looked up with the label of the first
    # atom of that residue number, libenergy.py:262-267)
    label_of_num, atoms_of_num = {}, {}
    for a in range(n):
        label_of_num.setdefault(int(resnums[a]), str(resnames[a]))
        atoms_of_num.setdefault(int(resnums[a]), []).append(a)
    for st in sysm.steps:
        if st.kind == "grow":
            _build_special_pairs(sysm, st, tables, label_of_num, atoms_of_num)
    _build_classes(sysm)
    _build_backbone_pairs(sysm, tables, label_of_num, atoms_of_num)


def build_system_from_labels(atom_labels, res_nums, res_labels, n_term, c_term, partial_indices=None,
                             terms=("lj", "clash", "coulomb"), tables=None):
    """GrowthSystem for the argument set of the reference's ``prepare_energy_function``
    (libenergy.py:27-61): per-atom labels of a partially grown structure and
    ``partial_indices = [new, old]``.  Supported index patterns are the two the reference's growth
    driver produces (side_chain_builder.py:81-85, 97-104): ``None`` (every atom scored against every
    other: the input-atom self energy) and ``[last n atoms of one residue, all atoms before them]``."""
    tables = tables or load_tables()
    names = np.array([str(x) for x in atom_labels]); resnames = np.array([str(x) for x in res_labels])
    resnums = np.array(res_nums, dtype=np.int64)
    n = len(names)
    steps = []
    if partial_indices is None:
        n_input = n
        appear = np.full(n, -1, dtype=np.int64)
    else:
        new_idx, old_idx = (np.asarray(x, dtype=np.int64) for x in partial_indices)
        n_new = len(new_idx)
        if not (np.array_equal(new_idx, np.arange(n - n_new, n)) and np.array_equal(old_idx, np.arange(n - n_new))):
            raise NotImplementedError("partial_indices must be [last n atoms, all earlier atoms] as built by "
                                      "initialize_func_calc (side_chain_builder.py:102-103)")
        num = int(resnums[new_idx[0]])
        if not (resnums[new_idx] == num).all():
            raise NotImplementedError("the new atoms must belong to one residue")
        label = str(resnames[np.where(resnums == num)[0][0]])          # backbone label of that residue
        tmpl = tables.templates[label]
        if [tmpl.names[k] for k in tmpl.sc_idx] != list(names[new_idx]):
            raise ValueError("new atoms are not the %s side chain in template order" % label)
        n_input = n - n_new
        appear = np.concatenate([np.full(n_input, -1), np.zeros(n_new)]).astype(np.int64)
        bb = [int(np.where((resnums[:n_input] == num) & (names[:n_input] == a))[0][0]) if
              ((resnums[:n_input] == num) & (names[:n_input] == a)).any() else 0 for a in ("N", "CA", "C")]
        steps.append(GrowthStep(index=0, resnum=num, resname=label, kind="grow", sc_start=n_input, n_sc=n_new,
                                bb_rows=tuple(bb)))
    sysm = GrowthSystem(
        names=names, resnames=resnames, resnums=resnums, chains=np.full(n, "A"), res_index=np.zeros(n, dtype=np.int64),
        is_nterm=np.asarray(n_term, dtype=bool).copy(), is_cterm=np.asarray(c_term, dtype=bool).copy(),
        charge=np.zeros(n), sigma=np.zeros(n), eps=np.zeros(n), radius=np.zeros(n),
        appear=appear, n_input=n_input, steps=steps, terms=tuple(terms))
    _finalize(sysm, tables)
    return sysm


def _build_special_pairs(sysm, st, tables, label_of_num, atoms_of_num):
    num = st.resnum
    new = list(range(st.sc_start, st.sc_start + st.n_sc))
    label = label_of_num[num]
    cand = []
    for r in (num - 1, num, num + 1):
        for b in atoms_of_num.get(r, ()):
            if sysm.appear[b] < st.index:          # present in the environment at this step
                cand.append(b)
    env, rels = [], {}
    for b in cand:
        rr = [relation(tables, sysm.names[a], num, label, sysm.names[b], int(sysm.resnums[b]),
                       label_of_num[int(sysm.resnums[b])]) for a in new]
        if int(sysm.resnums[b]) == num or any(r != REL_NONE for r in rr):
            env.append(b)
            rels[b] = rr
    ks, partners, coef = [], [], []
    P = lambda a: (sysm.charge[a], sysm.sigma[a], sysm.eps[a], sysm.radius[a])  # noqa: E731
    for ia, a in enumerate(new):                # new-new upper triangle, row-major (libcalc.py:62-69)
        for ib in range(ia + 1, len(new)):
            b = new[ib]
            rel = relation(tables, sysm.names[a], num, label, sysm.names[b], num, label)
            ks.append(ia); partners.append(-(ib + 1))
            coef.append(pair_coefficients(tables, rel, *P(a), *P(b), sysm.terms))
    for ia, a in enumerate(new):
        for je, b in enumerate(env):
            ks.append(ia); partners.append(je)
            coef.append(pair_coefficients(tables, rels[b][ia], *P(a), *P(b), sysm.terms))
    st.special_env = np.array(env, dtype=np.int64)
    st.sp_k = np.array(ks, dtype=np.int32)
    st.sp_partner = np.array(partners, dtype=np.int32)
    st.sp_coef = np.array(coef, dtype=np.float64).reshape(-1, 4)


def _build_classes(sysm):
    """Generic-pair classes = distinct (sigma, epsilon, clash element); slot layout sorted by
    (class, step of appearance) so the atoms present at any step are a prefix of each class."""
    key = [(sysm.sigma[a], sysm.eps[a], CLASH_ELEMENTS.index(sysm.names[a][0])) for a in range(sysm.n_atoms)]
    uniq = sorted(set(key))
    index = {k: i for i, k in enumerate(uniq)}
    sysm.cls_of_atom = np.array([index[k] for k in key], dtype=np.int32)
    sysm.n_cls = len(uniq)
    s = np.array([u[0] for u in uniq]); e = np.array([u[1] for u in uniq])
    r = np.array([VDW_RADII[CLASH_ELEMENTS[u[2]]] for u in uniq])
    use_lj, use_clash = "lj" in sysm.terms, "clash" in sysm.terms
    sysm.cls_s2 = ((s[:, None] + s[None, :]) * 5) ** 2
    sysm.cls_e4 = 4 * (e[:, None] * e[None, :]) ** 0.5 * (1.0 if use_lj else 0.0)
    sysm.cls_thr = (r[:, None] + r[None, :]) * CLASH_FACTOR * (1.0 if use_clash else 0.0)
    order = np.lexsort((np.arange(sysm.n_atoms), sysm.appear, sysm.cls_of_atom))
    sysm.atom_of_slot = order.astype(np.int32)
    sysm.slot_of_atom = np.empty(sysm.n_atoms, dtype=np.int32)
    sysm.slot_of_atom[order] = np.arange(sysm.n_atoms, dtype=np.int32)
    counts = np.bincount(sysm.cls_of_atom, minlength=sysm.n_cls)
    sysm.cls_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    L = sysm.n_res
    active = np.zeros((L + 1, sysm.n_cls), dtype=np.int32)
    for c in range(sysm.n_cls):
        ap = np.sort(sysm.appear[sysm.cls_of_atom == c])
        active[:, c] = np.searchsorted(ap, np.arange(L + 1), side="left")   # atoms with appear < i
    sysm.cls_active = active


def _build_backbone_pairs(sysm, tables, label_of_num, atoms_of_num):
    """Pairs of input atoms in the same or adjacent residue numbers, with exact coefficients;
    all other input-input pairs are generic.  Used for the backbone-only energy the growth
    starts from (side_chain_builder.py:149)."""
    ii, jj, coef = [], [], []
    P = lambda a: (sysm.charge[a], sysm.sigma[a], sysm.eps[a], sysm.radius[a])  # noqa: E731
    n_in = sysm.n_input
    for num, atoms in atoms_of_num.items():
        mine = [a for a in atoms if a < n_in]
        nxt = [a for a in atoms_of_num.get(num + 1, ()) if a < n_in]
        for x, a in enumerate(mine):
            for b in mine[x + 1:]:
                rel = relation(tables, sysm.names[a], num, label_of_num[num], sysm.names[b], num, label_of_num[num])
                ii.append(min(a, b)); jj.append(max(a, b))
                coef.append(pair_coefficients(tables, rel, *P(a), *P(b), sysm.terms))
            for b in nxt:
                rel = relation(tables, sysm.names[a], num, label_of_num[num], sysm.names[b], num + 1,
                               label_of_num[num + 1])
                ii.append(min(a, b)); jj.append(max(a, b))
                coef.append(pair_coefficients(tables, rel, *P(a), *P(b), sysm.terms))
    sysm.bb_pairs = (np.array(ii, dtype=np.int32), np.array(jj, dtype=np.int32),
                     np.array(coef, dtype=np.float64).reshape(-1, 4))


def placement_backbone(backbone_coords, sysm):
    """(L, 3, 3) float32 N,CA,C per residue as the reference sees them for template placement:
    read back from the 3-decimal structure table (side_chain_builder.py:143-144)."""
    rounded = _round3_f32(np.asarray(backbone_coords))
    return np.ascontiguousarray(rounded[_bb_rows(sysm)], dtype=np.float32)


def _bb_rows(sysm):
    """(L, 3) growth-order indices of every residue's N, CA, C (cached on the system)."""
    rows = getattr(sysm, "_bb_rows_arr", None)
    if rows is None:
        rows = np.array([st.bb_rows for st in sysm.steps], dtype=np.int64).reshape(sysm.n_res, 3)
        sysm._bb_rows_arr = rows
    return rows
