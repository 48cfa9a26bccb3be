"""CPU tests of the host tables the CUDA path consumes (mcsce_b200.system): the per-atom classes
plus the per-residue special-pair lists must reproduce, pair for pair, the coefficient arrays the
oracle's restatement of prepare_energy_function builds (and those are pinned to the reference by
tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from golden_utils import GoldenCase, library
from oracle import ref_numpy as O


def _system(name):
    from mcsce_b200.system import build_system

    case = GoldenCase(name)
    lib, ptm = library()
    s = case.structure()
    return case, s, build_system(s, retain_idxs=case.retain, terms=case.terms, rotamer_library=lib, ptm_library=ptm)


@pytest.mark.parametrize("name", ["pep12", "ptm14", "twochain", "pep6_ljclash", "fix12", "ptm18"])
def test_pair_coefficients_match_oracle(name):
    from mcsce_b200.chem.tables import load_tables

    case, s, sysm = _system(name)
    t = load_tables()
    work = s.copy()
    for st in sysm.steps:
        if st.kind == "skip":            # retained residue: its side chain came with the input structure
            continue
        tm = t.templates[st.resname]
        chain = sysm.chains[st.sc_start]
        work.append_atoms([tm.names[k] for k in tm.sc_idx], tm.resname, chain, st.resnum, np.zeros((tm.n_sc, 3)))
        if st.kind != "grow":
            continue
        n_all = len(work); n_sc = st.n_sc; n_old = n_all - n_sc
        assert n_old == st.sc_start
        n_term, c_term = work.get_terminal_res_atom_arr()
        ef = O.EnergyFunction(work.atom_labels, work.res_nums, work.res_labels, n_term, c_term,
                              np.arange(n_old, n_all), np.arange(n_old), terms=case.terms)
        nn = n_sc * (n_sc - 1) // 2
        A = np.nan_to_num(ef.acoeff, nan=0.0); B = np.nan_to_num(ef.bcoeff, nan=0.0)
        Q = np.nan_to_num(ef.charges, nan=0.0); thr = ef.thr
        if "lj" not in case.terms:
            A[:] = 0; B[:] = 0
        if "coulomb" not in case.terms:
            Q[:] = 0
        # rebuild the same arrays from classes + special list
        gA = np.zeros_like(A); gB = np.zeros_like(B); gQ = np.zeros_like(Q); gT = np.zeros_like(thr)
        q = sysm.charge if "coulomb" in case.terms else np.zeros_like(sysm.charge)
        for k in range(n_sc):
            a = st.sc_start + k
            ca = sysm.cls_of_atom[a]
            cb = sysm.cls_of_atom[:n_old]
            s2 = sysm.cls_s2[ca, cb]; e4 = sysm.cls_e4[ca, cb]
            sl = slice(nn + k * n_old, nn + (k + 1) * n_old)
            gA[sl] = e4 * s2 ** 6; gB[sl] = e4 * s2 ** 3
            gQ[sl] = q[a] * q[:n_old] * 0.25 * 1389.35
            gT[sl] = sysm.cls_thr[ca, cb]
        iu, ju = np.triu_indices(n_sc, k=1)
        tri = {(i, j): p for p, (i, j) in enumerate(zip(iu, ju))}
        seen_env = set()
        for k, pa, cf in zip(st.sp_k, st.sp_partner, st.sp_coef):
            if pa < 0:
                p = tri[(int(k), int(-pa - 1))]
            else:
                b = int(st.special_env[pa]); seen_env.add(b)
                p = nn + int(k) * n_old + b
            gA[p], gB[p], gQ[p], gT[p] = cf
        np.testing.assert_allclose(gA, A, rtol=2e-15, atol=0)
        np.testing.assert_allclose(gB, B, rtol=2e-15, atol=0)
        np.testing.assert_allclose(gQ, Q, rtol=2e-15, atol=0)
        np.testing.assert_allclose(gT, thr, rtol=1e-15, atol=0)
        # every environment atom with a bonded relation to the new side chain is in the special list
        rel_cols = np.where(np.isnan(ef.acoeff[nn:].reshape(n_sc, n_old)).any(0))[0]
        assert set(rel_cols) <= seen_env
        assert len(st.special_env) <= 16


def test_slot_layout_prefix_property():
    case, s, sysm = _system("pep12")
    n = sysm.n_atoms
    assert sorted(sysm.slot_of_atom) == list(range(n))
    for i in range(sysm.n_res + 1):
        present = set(np.where(sysm.appear < i)[0])
        got = set()
        for c in range(sysm.n_cls):
            lo = sysm.cls_start[c]
            got |= set(sysm.atom_of_slot[lo:lo + sysm.cls_active[i, c]])
        assert got == present


def test_backbone_pairs_match_oracle():
    """Input-atom self energy tables: generic classes + special list == oracle level-0 coefficients."""
    case, s, sysm = _system("twochain")
    n_term, c_term = s.get_terminal_res_atom_arr()
    ef = O.EnergyFunction(s.atom_labels, s.res_nums, s.res_labels, n_term, c_term, terms=case.terms)
    n = sysm.n_input
    iu, ju = np.triu_indices(n, k=1)
    ci, cj = sysm.cls_of_atom[iu], sysm.cls_of_atom[ju]
    gA = sysm.cls_e4[ci, cj] * sysm.cls_s2[ci, cj] ** 6
    gQ = sysm.charge[iu] * sysm.charge[ju] * 0.25 * 1389.35
    gT = sysm.cls_thr[ci, cj].copy()
    idx = {(i, j): p for p, (i, j) in enumerate(zip(iu, ju))}
    near = np.abs(sysm.resnums[iu] - sysm.resnums[ju]) <= 1
    bi, bj, coef = sysm.bb_pairs
    assert near.sum() == len(bi)
    for i, j, cf in zip(bi, bj, coef):
        p = idx[(int(i), int(j))]
        assert near[p]
        gA[p], gQ[p], gT[p] = cf[0], cf[2], cf[3]
    np.testing.assert_allclose(gA, np.nan_to_num(ef.acoeff), rtol=2e-15)
    np.testing.assert_allclose(gQ, np.nan_to_num(ef.charges), rtol=2e-15)
    np.testing.assert_allclose(gT, ef.thr, rtol=1e-15)


def test_placement_quaternions_reproduce_oracle_placement():
    from mcsce_b200.chem.tables import load_tables
    from mcsce_b200.libs.placement import _rotate, placement_quaternions

    t = load_tables()
    z = np.load(__import__("golden_utils").GOLDEN / "templates.npz")
    for res, tm in t.templates.items():
        bb = z[res + "_bb"]
        q1, q2, ca = placement_quaternions(bb, tm.coords[0], tm.coords[2])
        got = np.array([_rotate(q2, _rotate(q1, p).astype(np.float32)) + ca.astype(np.float64) for p in tm.coords])
        assert np.abs(got - z[res + "_placed"]).max() < 1e-9, res


def test_library_loads_and_exports_every_symbol():
    """The C-ABI library must exist in-tree and export everything include/mcsce_b200.h declares."""
    import re
    from pathlib import Path

    from mcsce_b200 import _lib

    lib = _lib.load_library()
    header = (Path(__file__).resolve().parents[1] / "include" / "mcsce_b200.h").read_text()
    declared = set(re.findall(r"\b(mcsce_[a-z_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.mcsce_version() >= 1


def test_engine_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from mcsce_b200.engine import GrowthEngine

    with pytest.raises(RuntimeError):
        GrowthEngine(0)


def test_trial_sets_follow_the_backbone_handed_in():
    """The reference derives phi/psi and with them the trial sets from the coordinates of every
    ``create_side_chain_structure`` call (side_chain_builder.py:143-147, 171): a sequence prepared on one conformation
    must give, for another conformation, the tables a fresh preparation on that conformation gives."""
    from mcsce_b200 import synth
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import assign_trials, build_system

    lib, ptm = library()
    seq = ["MET", "LYS", "PHE", "GLU", "LEU", "SER", "ARG", "ASP", "TRP", "GLN", "VAL", "ASN", "SEP", "ILE"]
    a = Structure(synth.small_peptide_pdb(seq, seed=2)).build().remove_side_chains([])                    # helix
    b = Structure(synth.small_peptide_pdb(seq, seed=2, phi=-120.0, psi=130.0)).build().remove_side_chains([])   # strand
    sys_a = build_system(a, rotamer_library=lib, ptm_library=ptm)
    sys_b = build_system(b, rotamer_library=lib, ptm_library=ptm)
    assert any(len(x.prob) != len(y.prob) or not np.array_equal(x.chi_mean, y.chi_mean)
               for x, y in zip(sys_a.steps, sys_b.steps) if x.kind == "grow")
    assert not assign_trials(sys_a, a.coords)             # same conformation: nothing to do
    assert assign_trials(sys_a, b.coords)
    for x, y in zip(sys_a.steps, sys_b.steps):
        assert (x.phi == y.phi or (np.isnan(x.phi) and np.isnan(y.phi))) and (x.psi == y.psi or (np.isnan(x.psi) and np.isnan(y.psi)))
        if x.kind == "grow":
            assert np.array_equal(x.chi_mean, y.chi_mean) and np.array_equal(x.chi_sigma, y.chi_sigma)
            assert np.array_equal(x.prob, y.prob) and x.n_chi == y.n_chi


def test_batched_placement_quaternions_are_bit_identical_to_the_per_residue_routine():
    from mcsce_b200.libs.placement import placement_quaternions, placement_quaternions_batch

    rng = np.random.default_rng(4)
    tn = np.array([-0.527, 1.359, 0.0], np.float32); tc = np.array([1.526, 0.0, 0.0], np.float32)
    for L in (1, 2, 3, 7, 16, 33, 257):
        bb = np.zeros((L, 3, 3), np.float32)
        for i in range(L):
            a = np.linalg.qr(rng.standard_normal((3, 3)))[0]
            bb[i] = np.round(np.stack([tn, np.zeros(3), tc]) @ a.T + rng.normal(0, 0.05, (3, 3)) + rng.uniform(-90, 90, 3), 3)
        TN = (np.tile(tn, (L, 1)) + rng.normal(0, 0.01, (L, 3))).astype(np.float32)
        TC = (np.tile(tc, (L, 1)) + rng.normal(0, 0.01, (L, 3))).astype(np.float32)
        q1, q2, ca = placement_quaternions_batch(bb, TN, TC)
        for i in range(L):
            r1, r2, rc = placement_quaternions(bb[i], TN[i], TC[i])
            assert np.array_equal(r1, q1[i]) and np.array_equal(r2, q2[i]) and np.array_equal(rc, ca[i]), (L, i)


def test_chains_that_restart_their_numbering_are_rejected():
    """Two chains numbered 1..6 each would alias onto the same residue ordinals (the reference maps ordinal <-> number as
    ordinal + res_nums[0], side_chain_builder.py:88-90); build_system must refuse instead of mis-assigning bonded pairs."""
    from mcsce_b200 import synth
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import build_system

    a = ["LYS", "GLU", "PHE", "SER", "TRP", "ASN"]
    s = Structure(synth.helix_bundle_pdb([a, a], seed=15, one_chain=False, spacing=12.0)).build().remove_side_chains([])
    build_system(s)                                           # consecutive numbering across chains: accepted
    num = s.cols["resseq"].copy()
    num[s.cols["chain"] != s.cols["chain"][0]] -= 6           # chain B restarts at the first number
    s.cols["resseq"] = num
    with pytest.raises(ValueError):
        build_system(s)


def test_synthetic_rotamer_library_is_opt_in(monkeypatch, tmp_path):
    """Production entry points must not silently sample from the seeded synthetic library (random priors): without
    $MCSCE_ROTAMER_LIB or a library under the package data directory the default path raises; the stand-in that tests and
    bench.py use needs MCSCE_B200_ALLOW_SYNTHETIC=1 and warns."""
    import warnings

    from mcsce_b200.core import rotamer_library as rl

    monkeypatch.delenv("MCSCE_ROTAMER_LIB", raising=False)
    monkeypatch.delenv(rl.ALLOW_SYNTHETIC_ENV, raising=False)
    with pytest.raises(FileNotFoundError):
        rl.default_library_path()
    monkeypatch.setenv(rl.ALLOW_SYNTHETIC_ENV, "1")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        p = rl.default_library_path()
    assert p.exists() and any("SYNTHETIC" in str(x.message) for x in w)
    monkeypatch.setenv("MCSCE_ROTAMER_LIB", str(tmp_path / "missing.lib"))
    with pytest.raises(FileNotFoundError):
        rl.default_library_path()
    real = tmp_path / "ALL.bbdep.rotamers.lib"
    real.write_text("# stand-in\n")
    monkeypatch.setenv("MCSCE_ROTAMER_LIB", str(real))
    assert rl.default_library_path() == real
