"""Pin the oracle (oracle/ref_numpy.py) to outputs of the running reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py from the
unmodified reference.  Tolerances: trial coordinates bit-exact (float32) up to the rare
last-bit difference of a float32 BLAS norm; per-trial energies 1e-11 relative (summation
order differs from numba's sequential nansum); clash flags and selected indices exact.
"""
import numpy as np
import pytest

from golden_utils import GoldenCase, library
from oracle import ref_numpy as O

FULL_CASES = ["pep6", "pep6_ljclash", "pep12", "ptm14", "twochain", "fix12", "ptm18"]
LIGHT_CASES = ["std20", "helix76", "bundle40"]


def _replay(case, c, record):
    lib, ptm = library()
    steps = case.steps_by_level(c)
    s = case.structure()

    def chi_for_step(i, res, phi, psi):
        _, _, prob = lib.trial_distribution(res, phi, psi, ptm)
        chis = steps[i]["chis"]
        assert len(prob) == len(chis), (i, res, len(prob), len(chis))
        return chis, prob

    def u_for_step(i):
        return float(steps[i]["u"])

    return O.grow_conformer(s, s.coords, case.beta, chi_for_step, u_for_step, retain_idxs=case.retain, terms=case.terms,
                            record=record)


def _check_case(name, max_conf=None, check_trial=True):
    case = GoldenCase(name)
    for c in range(case.n_conf if max_conf is None else min(max_conf, case.n_conf)):
        rec = []
        out = _replay(case, c, rec)
        gold = case.conf(c)
        steps = case.steps_by_level(c)
        assert out["ok"] == bool(gold["ok"])
        assert len(rec) == len(steps)
        for r in rec:
            g = steps[r["step"]]
            assert r["n_env"] == int(g["n_env"])
            ge, oe = g["energies"], r["energies"]
            assert np.array_equal(np.isinf(ge), np.isinf(oe)), "clash flags differ at step %d" % r["step"]
            fin = ~np.isinf(ge)
            np.testing.assert_allclose(oe[fin], ge[fin], rtol=1e-11, atol=1e-9)
            if check_trial and case.full:
                assert np.abs(r["trial"] - g["trial"]).max() <= 4e-6
                assert (r["trial"] != g["trial"]).mean() < 1e-3
            if "selected" in g:
                assert r["selected"] == int(g["selected"])
                np.testing.assert_allclose(r["weights"], g["weights"], rtol=1e-9, atol=1e-300)
        if out["ok"]:
            np.testing.assert_allclose(out["energy"], float(gold["energy"]), rtol=1e-11)
        bb_e = O.EnergyFunction  # noqa: F841  (backbone energy is part of out['energy'])


@pytest.mark.parametrize("name", FULL_CASES)
def test_growth_replay_full(name):
    _check_case(name)


@pytest.mark.parametrize("name", LIGHT_CASES)
def test_growth_replay_light(name):
    _check_case(name, check_trial=False)


@pytest.mark.parametrize("name,every", [("cfg4_150", 6), ("cfg3_300", 12)])
def test_growth_replay_full_size_sampled(name, every):
    """The stock reference at BASELINE sizes (150 residues with PTM residues / 300 residues): the NumPy oracle follows the
    recorded selections (one trial per residue to rebuild the environment) and re-scores every ``every``-th residue's
    whole trial set, first and last residue included - energies 1e-11, clash flags and the selection exact."""
    from mcsce_b200.chem.tables import load_tables
    from mcsce_b200.structure import _round3_f32

    case = GoldenCase(name)
    t = load_tables()
    s = case.structure()
    steps = case.steps_by_level(0)
    work = s.copy(); work.xyz = _round3_f32(np.asarray(s.coords, np.float32))
    ncac = work.get_sorted_minimal_backbone_coords()
    coords = np.asarray(s.coords, np.float32).copy()
    first = int(work.res_nums[0])
    grown = sorted(steps)
    sample = set(grown[::every]) | {grown[0], grown[-1]}
    lib, ptm = library()
    checked = 0
    for i, res in enumerate(work.residue_types):
        tm = t.templates[res]
        bb = ncac[3 * i:3 * i + 3]
        work.append_atoms([tm.names[k] for k in tm.sc_idx], tm.resname, work.residue_chains[i] if hasattr(work, "residue_chains") else "A",
                          first + i, np.zeros((tm.n_sc, 3), np.float32))
        if res in O.NO_CHI_RESIDUES:
            coords = np.concatenate([coords, O.place_template(bb, tm.coords)[tm.sc_idx].astype(np.float32)])
            continue
        g = steps[i]
        if i in sample:
            n_all = len(coords) + tm.n_sc
            n_term, c_term = work.get_terminal_res_atom_arr()
            ef = O.EnergyFunction(work.atom_labels, work.res_nums, work.res_labels, n_term, c_term,
                                  np.arange(n_all - tm.n_sc, n_all), np.arange(n_all - tm.n_sc), terms=case.terms, tables=t)
            trial = np.stack([O.trial_side_chain(tm, bb, ch) for ch in g["chis"]])
            e = ef.calculate(trial, np.broadcast_to(coords[None], (len(trial),) + coords.shape))
            ge = g["energies"]
            assert np.array_equal(np.isinf(e), np.isinf(ge)), (name, i)
            fin = ~np.isinf(ge)
            np.testing.assert_allclose(e[fin], ge[fin], rtol=1e-11, atol=1e-9)
            w = np.asarray(lib.trial_distribution(res, 0, 0, ptm)[2]) if False else None  # (weights are checked on the small fixtures)
            checked += 1
        coords = np.concatenate([coords, O.trial_side_chain(tm, bb, g["chis"][int(g["selected"])])])
    assert checked >= len(grown) // every


def test_final_coordinates_match_reference_table():
    """The reference returns coordinates rounded to 3 decimals in residue-number order."""
    from mcsce_b200.structure import _round3_f32

    case = GoldenCase("pep6")
    out = _replay(case, 0, [])
    gold = case.conf(0)
    s = case.structure()
    nums = list(s.res_nums)
    # growth order -> residue-number order (stable): side-chain blocks follow their backbone atoms
    from mcsce_b200.chem.tables import load_tables
    t = load_tables()
    first = nums[0]
    for i, res in enumerate(s.residue_types):
        nums.extend([first + i] * t.templates[res].n_sc)
    order = np.argsort(np.array(nums), kind="mergesort")
    got = _round3_f32(out["coords"])[order]
    assert np.abs(got - gold["final_coords"]).max() <= 1.1e-3
    assert (np.abs(got - gold["final_coords"]) > 1e-6).mean() < 0.01


def test_templates_chi_and_placement():
    from mcsce_b200.chem.tables import load_tables

    z = np.load(__import__("golden_utils").GOLDEN / "templates.npz")
    t = load_tables()
    for res, tm in t.templates.items():
        assert list(z[res + "_names"]) == tm.names
        assert np.array_equal(z[res + "_coords"], tm.coords)
        assert np.array_equal(z[res + "_sc_idx"], tm.sc_idx)
        if res in ("ALA", "GLY"):
            continue
        for n in range(1, 7):
            key = "%s_n%d_tors" % (res, n)
            if key not in z.files:
                break
            for tors, rot in zip(z[key], z["%s_n%d_rot" % (res, n)]):
                got = O.set_chis(tm, tors)
                assert np.abs(got - rot).max() <= 2e-6, (res, n)
        placed = O.place_template(z[res + "_bb"], tm.coords)
        assert np.abs(placed - z[res + "_placed"]).max() <= 1e-9, res
