"""A7 on the GPU: ``sasa_kernel`` / ``hpmf_pair_kernel`` through the C ABI (``mcsce_sasa``,
``mcsce_hpmf_energy``) against the fixtures the reference's own ``calc_sasa`` / ``init_hpmf_calculator``
produced, against the NumPy oracle at full size (config 3), and through size-independent properties.

Tolerance (north-star): |dV| <= max(1e-3, 1e-5 |V|); surface areas 1e-9 relative (float64 arithmetic on
bit-identical float32 distances; only the order of the product differs)."""
import numpy as np
import pytest

from golden_utils import GOLDEN, library

pytestmark = pytest.mark.gpu
Z = np.load(GOLDEN / "hpmf.npz", allow_pickle=False)
CASES = [str(c) for c in Z["cases"]]


def _case(name):
    return {k[len(name) + 1:]: Z[k] for k in Z.files if k.startswith(name + "_")}


def _tol(v):
    return max(1e-3, 1e-5 * abs(v))


@pytest.mark.parametrize("name", CASES)
def test_kernels_match_reference_functions(name):
    from mcsce_b200.hpmf import HpmfCalculator

    g = _case(name)
    calc = HpmfCalculator([str(t) for t in g["types"]], g["atom_filter"], (g["bond_i"], g["bond_j"]))
    v, sa = calc.energy(g["coords"], return_sasa=True)
    v, sa = v.cpu().numpy(), sa.cpu().numpy()
    np.testing.assert_allclose(sa, g["sasa"], rtol=1e-9, atol=1e-9)
    for k in range(len(v)):
        assert abs(v[k] - g["v"][k]) <= _tol(g["v"][k]), (k, v[k], g["v"][k])
    # the exposure decision of every carbon is the reference's
    carbon = calc.carbon
    assert np.array_equal(sa[:, carbon] > 6.0, g["sasa"][:, carbon] > 6.0)
    # stand-alone SASA entry point, one structure at a time = the batch
    one = calc.sasa(g["coords"][1]).cpu().numpy()
    assert np.array_equal(one, sa[1])
    assert np.array_equal(calc.sasa(g["coords"]).cpu().numpy(), sa)


def test_reference_shaped_closures():
    from mcsce_b200.libs.libhpmf import convert_atom_types, init_hpmf_calculator
    from mcsce_b200.libs.libsasa import calc_sasa

    g = _case("std20")
    names, resnums, labels = [str(x) for x in g["names"]], g["resnums"].tolist(), [str(x) for x in g["reslabels"]]
    atom_filter, types = convert_atom_types(names, resnums, labels)
    n = int(atom_filter.sum())
    conn = np.zeros((n, n)); conn[g["bond_i"], g["bond_j"]] = 1; conn += conn.T       # the reference's N x N matrix
    sasa_fn = calc_sasa(types, conn)
    hpmf_fn = init_hpmf_calculator(types, atom_filter, conn)
    x = g["coords"][0]
    np.testing.assert_allclose(sasa_fn(x[atom_filter]), g["sasa"][0], rtol=1e-9)
    v = hpmf_fn(x)
    assert isinstance(v, float) and abs(v - g["v"][0]) <= _tol(g["v"][0])
    vb = hpmf_fn(g["coords"])
    assert vb.shape == (5,) and vb[0] == v
    # a threshold nobody passes: no carbon pairs
    assert init_hpmf_calculator(types, atom_filter, conn, threshold=1e9)(x) == 0.0


def test_labels_path_equals_reference_connectivity():
    from mcsce_b200.hpmf import calculator_for_labels

    g = _case("bundle40")
    calc = calculator_for_labels([str(x) for x in g["names"]], g["resnums"].tolist(), [str(x) for x in g["reslabels"]])
    v = calc.energy(g["coords"]).cpu().numpy()
    for k in range(len(v)):
        assert abs(v[k] - g["v"][k]) <= _tol(g["v"][k])


def test_properties():
    """Size-independent checks: two copies of a structure far apart give twice the energy and the same
    surface areas; an isolated atom keeps its whole sphere; an atom order permutation changes nothing
    beyond summation order."""
    import torch

    from mcsce_b200.hpmf import HpmfCalculator, load_sasa_tables

    g = _case("helix76")
    types = [str(t) for t in g["types"]]
    f = g["atom_filter"]
    n, nt = len(f), int(f.sum())
    x = (np.round(g["coords"][0] * 1024) / 1024).astype(np.float32)    # multiples of 2^-10: x + 512 is exact in float32
    calc = HpmfCalculator(types, f, (g["bond_i"], g["bond_j"]))
    v1, sa1 = (t.cpu().numpy() for t in calc.energy(x, return_sasa=True))
    both = HpmfCalculator(types + types, np.concatenate([f, f]),
                          (np.concatenate([g["bond_i"], g["bond_i"] + nt]), np.concatenate([g["bond_j"], g["bond_j"] + nt])))
    x2 = np.concatenate([x, x + np.float32(512.0)])
    assert np.array_equal(x2[n:] - np.float32(512.0), x)
    v2, sa2 = (t.cpu().numpy() for t in both.energy(x2, return_sasa=True))
    assert abs(v2 - 2 * v1) <= 1e-9 * abs(v1)
    np.testing.assert_allclose(sa2[:nt], sa1, rtol=1e-12)
    # isolated atoms
    tb = load_sasa_tables()
    lone = HpmfCalculator(["CH3-SP3", "O-"], np.ones(2, bool), (np.zeros(0, int), np.zeros(0, int)))
    sa = lone.sasa(np.array([[0, 0, 0], [50, 0, 0]], np.float32)).cpu().numpy()
    np.testing.assert_allclose(sa, [4 * np.pi * (tb["R"][t] + 1.4) ** 2 for t in ("CH3-SP3", "O-")], rtol=1e-15)
    # permutation of the atoms
    rng = np.random.default_rng(0)
    perm = rng.permutation(n)
    inv = np.empty(n, int); inv[perm] = np.arange(n)
    fp = f[perm]
    typed_old = np.nonzero(f)[0]                      # typed ordinal -> atom
    new_typed_atoms = np.nonzero(fp)[0]               # new typed ordinal -> new atom position
    old_ord_of_atom = np.cumsum(f) - 1
    new_ord_of_old_ord = np.empty(nt, int)
    new_ord_of_old_ord[old_ord_of_atom[perm[new_typed_atoms]]] = np.arange(nt)
    types_p = [types[old_ord_of_atom[perm[a]]] for a in new_typed_atoms]
    pc = HpmfCalculator(types_p, fp, (new_ord_of_old_ord[g["bond_i"]], new_ord_of_old_ord[g["bond_j"]]))
    vp, sap = (t.cpu().numpy() for t in pc.energy(x[perm], return_sasa=True))
    np.testing.assert_allclose(sap[new_ord_of_old_ord], sa1, rtol=1e-12)
    assert abs(vp - v1) <= 1e-10 * abs(v1)
    assert typed_old.size == nt and torch.cuda.is_available()


def test_config3_grown_conformers_against_oracle():
    """HPMF of freshly grown 300-residue conformers (4907 atoms): the batch on the GPU against the
    NumPy oracle on three of them; deterministic; dead conformers (NaN coordinates never enter)."""
    from mcsce_b200 import synth
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.hpmf import calculator_for_labels, convert_atom_types, bond_pairs, load_sasa_tables
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import build_system
    from oracle import hpmf_numpy as O

    lib, ptm = library()
    s = Structure(synth.config_backbone(3)).build().remove_side_chains([])
    sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    out = eng.grow(256, 1.0 / (300 * 0.008314462), seed=5)
    ok = out["ok"].cpu().numpy().astype(bool)
    coords = out["coords"][out["ok"].bool()]
    assert ok.sum() > 100
    names, resnums, labels = [str(a) for a in sysm.names], [int(r) for r in sysm.resnums], [str(r) for r in sysm.resnames]
    calc = calculator_for_labels(names, resnums, labels)
    v, sa = calc.energy(coords, return_sasa=True)
    v2 = calc.energy(coords)
    assert np.array_equal(v.cpu().numpy(), v2.cpu().numpy())
    v, sa, xs = v.cpu().numpy(), sa.cpu().numpy(), coords.cpu().numpy()
    assert np.isfinite(v).all() and (v < 0).all() and len(np.unique(v)) == len(v)
    keep, types = convert_atom_types(names, resnums, labels)
    bi, bj = bond_pairs(names, resnums, labels)
    idx = np.cumsum(keep) - 1
    sel = keep[bi] & keep[bj]
    bonded = np.zeros((keep.sum(), keep.sum()), bool)
    bonded[idx[bi[sel]], idx[bj[sel]]] = True
    bonded |= bonded.T
    for c in (0, len(v) // 2, len(v) - 1):
        vo, sao = O.hpmf(xs[c], keep, types, bonded, load_sasa_tables())
        np.testing.assert_allclose(sa[c], sao, rtol=1e-9, atol=1e-9)
        assert abs(v[c] - vo) <= _tol(vo), (c, v[c], vo)


# ---------------------------------------------------------------------------------------------
# the term inside the growth (step-by-step loop through the single-kernel entry points)
# ---------------------------------------------------------------------------------------------

def _oracle_hpmf_term():
    from mcsce_b200.hpmf import bond_pairs, load_sasa_tables
    from oracle import hpmf_numpy as O

    cache = {}

    def term(labels, nums, reslabels, coords):
        n = len(labels)
        if n not in cache:
            names, rn, rl = [str(a) for a in labels], [int(x) for x in nums], [str(x) for x in reslabels]
            keep, types = O.convert_atom_types(names, rn, rl, load_sasa_tables()["atom_types"])
            bi, bj = bond_pairs(names, rn, rl)
            idx = np.cumsum(keep) - 1
            sel = keep[bi] & keep[bj]
            bonded = np.zeros((keep.sum(), keep.sum()), bool)
            bonded[idx[bi[sel]], idx[bj[sel]]] = True
            cache[n] = (keep, types, bonded | bonded.T)
        keep, types, bonded = cache[n]
        return O.hpmf(np.asarray(coords, np.float32), keep, types, bonded, load_sasa_tables())[0]

    return term


def _pep12():
    from golden_utils import GoldenCase
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.system import build_system

    case = GoldenCase("pep12")
    lib, ptm = library()
    s = case.structure()
    sysm = build_system(s, terms=case.terms, rotamer_library=lib, ptm_library=ptm)
    return case, s, sysm, GrowthEngine(0).set_system(sysm).set_backbone(s.coords)


def test_stepwise_growth_equals_fused_growth():
    """Same trial rows and uniforms through ``mcsce_grow`` and through the host-driven K1/K2/K3 loop: same
    selections, same coordinates, energies within the per-trial tolerance summed over the residues."""
    case, s, sysm, eng = _pep12()
    u = np.random.default_rng(3).random((8, sysm.n_res))
    a = eng.grow(8, case.beta, seed=11, uniforms=u, dump=True)
    chis = a["chis"].cpu().numpy()
    b = eng.grow_stepwise(8, case.beta, chis=chis, uniforms=u)
    sel_a, sel_b = a["selected"].cpu().numpy(), b["selected"].cpu().numpy()
    grown = [st.index for st in sysm.steps if st.kind == "grow"]
    ok = a["ok"].cpu().numpy().astype(bool)
    assert np.array_equal(sel_a[ok][:, grown], sel_b[ok][:, grown])
    assert np.array_equal(ok, b["ok"].cpu().numpy().astype(bool)) and ok.sum() >= 4
    assert np.array_equal(a["coords"].cpu().numpy()[ok], b["coords"].cpu().numpy()[ok])
    ea, eb = a["energy"].cpu().numpy()[ok], b["energy"].cpu().numpy()[ok]
    assert (np.abs(ea - eb) <= 1e-3 * len(grown) + 1e-5 * np.abs(ea)).all()
    # own generator: deterministic in the seed, different between seeds
    c1 = eng.grow_stepwise(4, case.beta, seed=5)
    c2 = eng.grow_stepwise(4, case.beta, seed=5)
    c3 = eng.grow_stepwise(4, case.beta, seed=6)
    assert np.array_equal(c1["coords"].cpu().numpy(), c2["coords"].cpu().numpy())
    assert not np.array_equal(c1["selected"].cpu().numpy(), c3["selected"].cpu().numpy())


def test_growth_with_hpmf_term_against_oracle_growth():
    """lj + clash + coulomb + hpmf: the GPU's step-by-step growth against the NumPy oracle's growth loop carrying
    the same coordinate-based term, on identical trial rows and uniforms - per-trial energies within tolerance,
    every selection identical."""
    from mcsce_b200.hpmf import growth_term
    from oracle import ref_numpy as O

    case, s, sysm, eng = _pep12()
    term = growth_term(eng)
    u = np.random.default_rng(4).random((3, sysm.n_res))
    rec, rec_bf = [], []
    out = eng.grow_stepwise(3, case.beta, seed=9, uniforms=u, extra_term=term, record=rec)
    # the brute-force form of the term (one whole structure per trial) gives the same growth
    bf = eng.grow_stepwise(3, case.beta, seed=9, uniforms=u, extra_term=growth_term(eng, term.calculator, brute_force=True),
                           record=rec_bf)
    assert np.array_equal(bf["selected"].cpu().numpy(), out["selected"].cpu().numpy())
    for a, b in zip(rec, rec_bf):
        fin = np.isfinite(a["energies"])
        assert np.array_equal(fin, np.isfinite(b["energies"]))
        assert np.allclose(a["energies"][fin], b["energies"][fin], rtol=1e-11, atol=1e-9)
    plain = eng.grow_stepwise(3, case.beta, seed=9, uniforms=u)
    sel = out["selected"].cpu().numpy()
    assert not np.array_equal(sel, plain["selected"].cpu().numpy())          # the term changes the growth
    by_step = {r["step"]: r for r in rec}
    oterm = _oracle_hpmf_term()
    worst, outliers, n_trials = 0.0, 0, 0
    for c in range(3):
        def chi_for_step(i, res, phi, psi):
            st = sysm.steps[i]
            return by_step[i]["chis"][c, :, :st.n_chi_lib], st.prob

        orec = []
        ref = O.grow_conformer(s, s.coords, case.beta, chi_for_step, lambda i: u[c, i], terms=case.terms, record=orec,
                               extra_term=oterm)
        assert bool(out["ok"][c].item()) == ref["ok"]
        for r in orec:
            got = by_step[r["step"]]["energies"][c]
            assert np.array_equal(np.isinf(got), np.isinf(r["energies"])), r["step"]
            fin = np.isfinite(r["energies"])
            tol = np.maximum(1e-3, 1e-5 * np.abs(r["energies"][fin]))
            ratio = np.abs(got[fin] - r["energies"][fin]) / tol
            # a carbon whose surface area sits on the 6 A^2 exposure threshold switches all its pair terms with a
            # 1e-6 A coordinate difference (K1 fidelity): such trials are counted, not bounded
            outliers += int((ratio > 4.0).sum()); n_trials += int(fin.sum())
            if (ratio <= 4.0).any():
                worst = max(worst, float(ratio[ratio <= 4.0].max()))
            if "selected" in r:
                assert sel[c, r["step"]] == r["selected"], (c, r["step"])
        if ref["ok"]:
            e = float(out["energy"][c].item())
            assert abs(e - ref["energy"]) <= 1e-3 * sysm.n_res + 1e-5 * abs(ref["energy"]), (e, ref["energy"])
    assert n_trials > 500 and outliers <= 2, (outliers, n_trials, worst)


def test_hpmf_through_the_reference_entry_points(tmp_path):
    """terms=[..., "hpmf"] through prepare_energy_function's closure and initialize_func_calc /
    create_side_chain_ensemble."""
    from functools import partial

    from mcsce_b200.core import side_chain_builder as scb
    from mcsce_b200.libs.libenergy import prepare_energy_function
    from oracle import ref_numpy as O

    case, s, sysm, eng = _pep12()
    terms = ["lj", "clash", "coulomb", "hpmf"]
    oterm = _oracle_hpmf_term()
    # closure of a level: residue 1 (ARG) grown on top of MET
    out = eng.grow_stepwise(1, case.beta, seed=2)
    xyz = out["coords"][0].cpu().numpy()
    st = [x for x in sysm.steps if x.kind == "grow"][1]
    n_all = st.sc_start + st.n_sc
    names, nums, labels = sysm.names[:n_all], sysm.resnums[:n_all], sysm.resnames[:n_all]
    idx = np.arange(n_all)
    f = prepare_energy_function(names, nums, labels, sysm.is_nterm[:n_all], sysm.is_cterm[:n_all], None,
                                partial_indices=[idx[-st.n_sc:], idx[:-st.n_sc]], terms=terms)
    f_plain = prepare_energy_function(names, nums, labels, sysm.is_nterm[:n_all], sysm.is_cterm[:n_all], None,
                                      partial_indices=[idx[-st.n_sc:], idx[:-st.n_sc]], terms=terms[:3])
    rows = np.zeros((6, 6)); rows[:, :st.n_chi_lib] = st.chi_mean[:6]
    trial = eng.place_trials(st.index, rows).cpu().numpy()
    env = np.repeat(xyz[None, :st.sc_start], 6, axis=0)
    got, base = f(trial, env), f_plain(trial, env)
    assert np.array_equal(np.isinf(got), np.isinf(base))
    for k in np.nonzero(np.isfinite(base))[0]:
        v = oterm(names, nums, labels, np.concatenate([env[k], trial[k]]))
        assert abs((got[k] - base[k]) - v) <= max(1e-3, 1e-5 * abs(v)), (k, got[k] - base[k], v)
    # whole-structure closure (level 0)
    n_term, c_term = s.get_terminal_res_atom_arr()
    f0 = prepare_energy_function(s.atom_labels, s.res_nums, s.res_labels, n_term, c_term, None, terms=terms)
    f0p = prepare_energy_function(s.atom_labels, s.res_nums, s.res_labels, n_term, c_term, None, terms=terms[:3])
    v0 = oterm(s.atom_labels, s.res_nums, s.res_labels, s.coords)
    assert abs((f0(s.coords[None], s.coords[None, :0])[0] - f0p(s.coords[None], s.coords[None, :0])[0]) - v0) <= 1e-3
    # growth entry points
    scb.set_seed(3)
    scb.initialize_func_calc(partial(prepare_energy_function, batch_size=16, forcefield=None, terms=terms), structure=s)
    confs, success = scb.create_side_chain_ensemble(s, 6, 300, str(tmp_path))
    assert len(success) == 6 and sum(success) >= 3
    lines = (tmp_path / "energies.csv").read_text().strip().splitlines()
    assert len(lines) == 1 + sum(success)
    structure, ok, energy, _ = scb.create_side_chain_structure([s.coords, case.beta, [], None])
    assert ok and np.isfinite(energy)
    with pytest.raises(NotImplementedError):
        scb.initialize_func_calc(partial(prepare_energy_function, terms=["lj", "gb"]), structure=s)
    # leave the module prepared for the plain terms (other tests share it)
    scb.initialize_func_calc(partial(prepare_energy_function, batch_size=16, forcefield=None, terms=terms[:3]), structure=s)


@pytest.mark.parametrize("cfg,n_conf", [(2, 6), (3, 3)])
def test_per_trial_increments_equal_whole_structure_evaluation(cfg, n_conf):
    """``mcsce_hpmf_trials`` (environment once, every trial adds its own factors and pair terms) against one
    whole-structure ``mcsce_hpmf_energy`` per trial, on partially grown conformers of configs 2 and 3 at an early,
    a middle and a late residue, with and without a mask."""
    import torch

    from mcsce_b200 import synth
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.hpmf import calculator_for_labels
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import build_system

    lib, ptm = library()
    s = Structure(synth.config_backbone(cfg)).build().remove_side_chains([])
    sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    out = eng.grow(4 * n_conf, 1.0 / (300 * 0.008314462), seed=8)
    full = out["coords"][out["ok"].bool()][:n_conf].contiguous()
    C_ = full.shape[0]
    assert C_ >= 2
    calc = calculator_for_labels([str(a) for a in sysm.names], [int(r) for r in sysm.resnums], [str(r) for r in sysm.resnames])
    grown = [st for st in sysm.steps if st.kind == "grow"]
    rng = np.random.default_rng(1)
    worst = 0.0
    distinct = []
    for st in (grown[1], grown[len(grown) // 2], grown[-1]):
        T = min(len(st.prob), 24)
        rows = np.zeros((C_ * T, 6))
        rows[:, :st.n_chi_lib] = np.tile(st.chi_mean[:T], (C_, 1)) + rng.normal(0, 5.0, (C_ * T, st.n_chi_lib))
        xyz = eng.place_trials(st.index, rows).reshape(C_, T, st.n_sc, 3).contiguous()
        got = calc.trial_energies(full, st.sc_start, st.n_sc, xyz).cpu().numpy()
        end = st.sc_start + st.n_sc
        structs = full[:, None].repeat(1, T, 1, 1)
        structs[:, :, st.sc_start:end] = xyz
        structs[:, :, end:] = 1.0e15
        ref = calc.energy(structs.reshape(C_ * T, -1, 3)).cpu().numpy().reshape(C_, T)
        err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
        worst = max(worst, float(err.max()))
        assert err.max() <= 1e-10, (st.index, st.resname, err.max())
        # trials and conformers differ (a side chain whose only carbon is CB - tanh(SA) saturated, position independent of
        # chi - leaves V the same for every trial to 1e-9: required of the conformers always, of the trials at one residue)
        assert len(np.unique(np.round(ref, 9))) >= C_
        distinct.append(len(np.unique(np.round(ref, 9))) > T)
        mask = torch.as_tensor(rng.random((C_, T)) < 0.5, device=full.device)
        got_m = calc.trial_energies(full, st.sc_start, st.n_sc, xyz, mask).cpu().numpy()
        mk = mask.cpu().numpy()
        assert np.array_equal(got_m[mk], got[mk]) and (got_m[~mk] == 0).all()
    assert any(distinct)
    print("worst relative difference", worst)


@pytest.mark.parametrize("cfg,n_conf", [(2, 12), (3, 6)])
def test_carried_environment_state_equals_rederiving_it(cfg, n_conf):
    """The hpmf growth with the environment's areas / weights / fields extended from residue to residue
    (``mcsce_hpmf_carry``) against the same growth re-deriving them at every residue: identical selections,
    per-trial energies to 1e-10 relative, incl. conformers that die on the way."""
    from mcsce_b200 import synth
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.hpmf import growth_term
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import build_system

    lib, ptm = library()
    s = Structure(synth.config_backbone(cfg)).build().remove_side_chains([])
    sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    beta = 1.0 / (300 * 0.008314462)
    on = growth_term(eng, carry=True)
    off = growth_term(eng, on.calculator, carry=False)
    ra, rb = [], []
    a = eng.grow_stepwise(n_conf, beta, seed=13, extra_term=on, record=ra)
    b = eng.grow_stepwise(n_conf, beta, seed=13, extra_term=off, record=rb)
    assert np.array_equal(a["selected"].cpu().numpy(), b["selected"].cpu().numpy())
    assert np.array_equal(a["ok"].cpu().numpy(), b["ok"].cpu().numpy())
    worst = 0.0
    for x, y in zip(ra, rb):
        fin = np.isfinite(y["energies"])
        assert np.array_equal(fin, np.isfinite(x["energies"]))
        if fin.any():
            worst = max(worst, float((np.abs(x["energies"][fin] - y["energies"][fin]) / np.maximum(1, np.abs(y["energies"][fin]))).max()))
    assert worst <= 1e-10, worst
    ok = a["ok"].cpu().numpy().astype(bool)
    np.testing.assert_allclose(a["energy"].cpu().numpy()[ok], b["energy"].cpu().numpy()[ok], rtol=1e-11)
    # a second growth on the same calculator starts from scratch (state forgotten)
    a2 = eng.grow_stepwise(n_conf, beta, seed=13, extra_term=on)
    assert np.array_equal(a2["selected"].cpu().numpy(), a["selected"].cpu().numpy())
    print("carried vs re-derived, worst relative difference", worst)


def _workload(cfg):
    from mcsce_b200 import synth
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import build_system

    lib, ptm = library()
    s = Structure(synth.config_backbone(cfg)).build().remove_side_chains([])
    sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
    return s, sysm, GrowthEngine(0).set_system(sysm).set_backbone(s.coords)


def _calc_for(sysm, eng):
    from mcsce_b200.hpmf import calculator_for_labels

    return calculator_for_labels([str(a) for a in sysm.names], [int(r) for r in sysm.resnums],
                                 [str(r) for r in sysm.resnames], device=eng.device_index)


@pytest.mark.parametrize("which,n_conf", [("pep12", 3), ("cfg2", 6), ("cfg3", 4)])
def test_fused_growth_with_hpmf_equals_the_host_driven_loop(which, n_conf):
    """terms = [lj, clash, coulomb, hpmf] inside ``mcsce_grow`` (mcsce_grow_opts.hpmf: Philox streams, the residue's near
    lists, the environment state carried on the device) against the host-driven loop that calls the single-kernel entry
    points with the term evaluated over the whole environment - the loop the oracle comparison above pins: same chi rows
    and uniforms -> every selection identical, per-trial energies equal to summation order."""
    from mcsce_b200.hpmf import growth_term

    if which == "pep12":
        case, s, sysm, eng = _pep12()
        beta = case.beta
    else:
        s, sysm, eng = _workload(int(which[3:]))
        beta = 1.0 / (300 * 0.008314462)
    calc = _calc_for(sysm, eng)
    u = np.random.default_rng(8).random((n_conf, sysm.n_res))
    fused = eng.grow(n_conf, beta, seed=21, uniforms=u, dump=True, hpmf=calc)
    plain = eng.grow(n_conf, beta, seed=21, uniforms=u, dump=True)
    chis = fused["chis"].cpu().numpy()
    pch = plain["chis"].cpu().numpy()                                      # the term does not touch the trial rows
    both = (chis != 0) & (pch != 0)                                        # (rows past a failed residue are written or not, by driver)
    assert np.array_equal(chis[both], pch[both]) and both.sum() > 0.9 * (chis != 0).sum()
    rec = []
    step = eng.grow_stepwise(n_conf, beta, chis=chis, uniforms=u, extra_term=growth_term(eng, calc), record=rec)
    sel_f, sel_s = fused["selected"].cpu().numpy(), step["selected"].cpu().numpy()
    grown = [st.index for st in sysm.steps if st.kind == "grow"]
    ok_f = fused["ok"].cpu().numpy()
    assert (ok_f != 2).all()
    assert np.array_equal(ok_f.astype(bool), step["ok"].cpu().numpy().astype(bool))
    te = fused["trial_energy"].cpu().numpy()
    n_cmp = 0
    for c in range(n_conf):
        for r in rec:
            i = r["step"]
            assert sel_f[c, i] == sel_s[c, i], (which, c, i)
            if sel_f[c, i] < 0:
                break
            o = eng.trial_off[i]
            got, ref = te[c, o:o + r["energies"].shape[1]], r["energies"][c]
            assert np.array_equal(np.isinf(got), np.isinf(ref))
            fin = np.isfinite(ref)
            # (the lj/coulomb part differs at float32-sum level between the two loops: the fused one scores the chi-independent
            #  atoms once per residue, the single-kernel entry point scores them in every trial; the hpmf part is the same
            #  arithmetic - with and without the residue lists it agrees to 2e-13, tools/diag_hpmf_lists.py.  The float32 error
            #  scales with the size of the partial sums, not of the total: half the north-star absolute tolerance is allowed)
            assert np.allclose(got[fin], ref[fin], rtol=2e-6, atol=5e-4), (which, c, i, np.abs(got[fin] - ref[fin]).max())
            n_cmp += int(fin.sum())
    assert n_cmp > 100
    ok = ok_f.astype(bool)
    assert np.array_equal(fused["coords"].cpu().numpy()[ok], step["coords"].cpu().numpy()[ok])
    ef, es = fused["energy"].cpu().numpy()[ok], step["energy"].cpu().numpy()[ok]
    assert np.allclose(ef, es, rtol=2e-6, atol=1e-4)
    # the term changes the growth
    assert not np.array_equal(sel_f[:, grown], plain["selected"].cpu().numpy()[:, grown])


def test_fused_hpmf_growth_in_two_conformer_groups():
    """>= 192 conformers grow as two groups on separate streams, each with its own carried environment state."""
    s, sysm, eng = _workload(2)
    calc = _calc_for(sysm, eng)
    beta = 1.0 / (300 * 0.008314462)
    n, half = 256, 128
    whole = eng.grow(n, beta, seed=5, hpmf=calc, env_split=1, host_out=True)
    e, ok, xyz = whole["energy"].copy(), whole["ok"].copy(), whole["coords"].copy()
    assert (ok == 1).mean() > 0.5 and (ok != 2).all()
    for lo in (0, half):
        part = eng.grow(half, beta, seed=5, conf_id0=lo, hpmf=calc, env_split=1, host_out=True)
        assert np.array_equal(part["ok"], ok[lo:lo + half])
        assert np.array_equal(part["energy"], e[lo:lo + half], equal_nan=True)
        assert np.array_equal(part["coords"], xyz[lo:lo + half])
    again = eng.grow(n, beta, seed=5, hpmf=calc, env_split=1, host_out=True)
    assert np.array_equal(again["energy"], e, equal_nan=True)


def test_per_conformer_trial_kernel_equals_per_trial_kernel():
    """Inside mcsce_grow the trials of a conformer are evaluated by one block (a warp per trial, the residue's lists staged in
    shared memory); ``mcsce_hpmf_set_option(h, 0, 0)`` switches the growth back to the block-per-trial kernel that
    ``mcsce_hpmf_trials`` runs (pinned against whole-structure evaluations above): same selections, energies to 1e-12."""
    s, sysm, eng = _workload(3)
    calc = _calc_for(sysm, eng)
    beta = 1.0 / (300 * 0.008314462)
    u = np.random.default_rng(2).random((6, sysm.n_res))
    a = eng.grow(6, beta, seed=4, uniforms=u, dump=True, hpmf=calc)
    assert calc.lib.mcsce_hpmf_set_option(calc.h, 0, 0) == 0
    b = eng.grow(6, beta, seed=4, uniforms=u, dump=True, hpmf=calc)
    assert calc.lib.mcsce_hpmf_set_option(calc.h, 0, 1) == 0
    assert np.array_equal(a["selected"].cpu().numpy(), b["selected"].cpu().numpy())
    ta, tb = a["trial_energy"].cpu().numpy(), b["trial_energy"].cpu().numpy()
    assert np.array_equal(np.isfinite(ta), np.isfinite(tb))
    fin = np.isfinite(ta)
    assert fin.sum() > 10000 and np.allclose(ta[fin], tb[fin], rtol=1e-12, atol=1e-10)
