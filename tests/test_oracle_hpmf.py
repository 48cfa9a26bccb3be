"""A7 (hydrophobic PMF) on the CPU: the NumPy oracle against the fixtures made by the reference's own
``calc_sasa`` / ``init_hpmf_calculator`` (tests/golden/make_golden_hpmf.py), and the host tables
(atom typing, bond connectivity) against what the reference produced."""
import numpy as np
import pytest

from golden_utils import GOLDEN
from oracle import hpmf_numpy as O

Z = np.load(GOLDEN / "hpmf.npz", allow_pickle=False)
CASES = [str(c) for c in Z["cases"]]


def _case(name):
    g = {k[len(name) + 1:]: Z[k] for k in Z.files if k.startswith(name + "_")}
    n = int(g["atom_filter"].sum())
    bonded = np.zeros((n, n), dtype=bool)
    bonded[g["bond_i"], g["bond_j"]] = True
    bonded |= bonded.T
    g["bonded"] = bonded
    return g


@pytest.mark.parametrize("name", CASES)
def test_oracle_sasa_and_hpmf_match_reference_functions(name):
    from mcsce_b200.hpmf import load_sasa_tables

    tb = load_sasa_tables()
    g = _case(name)
    types = [str(t) for t in g["types"]]
    for k in range(len(g["coords"])):
        v, sa = O.hpmf(g["coords"][k], g["atom_filter"], types, g["bonded"], tb)
        np.testing.assert_allclose(sa, g["sasa"][k], rtol=1e-12, atol=1e-12)
        assert abs(v - g["v"][k]) <= 1e-9 * max(1.0, abs(g["v"][k])), (k, v, g["v"][k])


@pytest.mark.parametrize("name", CASES)
def test_host_typing_and_bonds_match_reference(name):
    from mcsce_b200.hpmf import bond_pairs, convert_atom_types

    g = _case(name)
    names, resnums, labels = [str(x) for x in g["names"]], g["resnums"].tolist(), [str(x) for x in g["reslabels"]]
    keep, types = convert_atom_types(names, resnums, labels)
    assert (keep == g["atom_filter"]).all()
    assert types == [str(t) for t in g["types"]]
    # oracle's restatement of the typing too
    from mcsce_b200.hpmf import load_sasa_tables
    keep_o, types_o = O.convert_atom_types(names, resnums, labels, load_sasa_tables()["atom_types"])
    assert (keep_o == keep).all() and types_o == types
    bi, bj = bond_pairs(names, resnums, labels)
    idx = np.cumsum(keep) - 1
    sel = keep[bi] & keep[bj]
    mine = set(zip(idx[bi[sel]].tolist(), idx[bj[sel]].tolist()))
    ref = set(zip(g["bond_i"].tolist(), g["bond_j"].tolist()))
    assert mine == ref


def test_unknown_residue_raises_like_reference():
    from mcsce_b200.hpmf import convert_atom_types

    with pytest.raises(KeyError):                 # the reference's type table has no PTM residues
        convert_atom_types(["N", "CA"], [2, 2], ["SEP", "SEP"])


def test_bonded_table_limits():
    from mcsce_b200.hpmf import _bonded_table

    t = _bonded_table(3, (np.array([0, 1]), np.array([1, 2])))
    assert t.tolist() == [[1, -1, -1, -1], [0, 2, -1, -1], [1, -1, -1, -1]]
    m = np.zeros((6, 6)); m[0, 1:] = 1; m += m.T
    with pytest.raises(ValueError):
        _bonded_table(6, m)


def test_oracle_properties():
    """Size-independent properties of the restatement itself: atom order does not matter, two copies far apart
    give twice the energy, an isolated atom keeps its whole sphere, a huge threshold leaves no pairs."""
    from mcsce_b200.hpmf import load_sasa_tables

    tb = load_sasa_tables()
    g = _case("pep12")
    types = [str(t) for t in g["types"]]
    f = g["atom_filter"]
    x = g["coords"][0]
    v, sa = O.hpmf(x, f, types, g["bonded"], tb)
    # permutation of the typed atoms
    rng = np.random.default_rng(0)
    xt = x[f]
    perm = rng.permutation(len(types))
    vp, sap = O.hpmf(xt[perm], np.ones(len(types), bool), [types[k] for k in perm], g["bonded"][np.ix_(perm, perm)], tb)
    np.testing.assert_allclose(sap, sa[perm], rtol=1e-12)
    assert abs(vp - v) <= 1e-10 * abs(v)
    # two copies 512 A apart (coordinates on a 2^-10 grid so that the shift is exact in float32)
    xq = (np.round(xt * 1024) / 1024).astype(np.float32)
    v1, _ = O.hpmf(xq, np.ones(len(types), bool), types, g["bonded"], tb)
    n = len(types)
    both = np.zeros((2 * n, 2 * n), bool); both[:n, :n] = g["bonded"]; both[n:, n:] = g["bonded"]
    v2, _ = O.hpmf(np.concatenate([xq, xq + np.float32(512)]), np.ones(2 * n, bool), types + types, both, tb)
    assert abs(v2 - 2 * v1) <= 1e-10 * abs(v1)
    lone = O.sasa(np.zeros((1, 3), np.float32), ["CH3-SP3"], np.zeros((1, 1), bool), tb["P"], tb["R"], tb["pij_bonded"],
                  tb["pij_non_bonded"])
    assert lone[0] == 4 * np.pi * (tb["R"]["CH3-SP3"] + 1.4) ** 2
    assert O.hpmf(x, f, types, g["bonded"], tb, threshold=1e9)[0] == 0.0


def test_oracle_growth_hook_is_neutral_when_the_term_is_zero():
    """``extra_term`` of the oracle's growth loop: a zero term reproduces the fixture's growth, a constant one only
    shifts the accumulated energy by (grown residues + 1) x constant."""
    from golden_utils import GoldenCase, library
    from oracle import ref_numpy as R

    lib, ptm = library()
    case = GoldenCase("pep6")
    s = case.structure()
    steps = {k: case.step(0, k) for k in range(case.n_steps(0))}
    by_level = case.steps_by_level(0)

    def chi_for_step(i, res, phi, psi):
        return by_level[i]["chis"], lib.trial_distribution(res, phi, psi, ptm)[2]

    runs = []
    for const in (None, 0.0, 2.5):
        term = None if const is None else (lambda labels, nums, rl, xyz, c=const: c)
        rec = []
        out = R.grow_conformer(s, s.coords, case.beta, chi_for_step, lambda i: float(by_level[i]["u"]), terms=case.terms,
                               record=rec, extra_term=term)
        runs.append((out, rec))
    (a, ra), (b, rb), (c, rc) = runs
    assert a["ok"] and b["ok"] and c["ok"] and len(steps) > 0
    assert a["energy"] == b["energy"] and np.array_equal(a["coords"], b["coords"])
    assert [r["selected"] for r in ra] == [r["selected"] for r in rc]
    assert abs(c["energy"] - a["energy"] - 2.5 * (len(rc) + 1)) <= 1e-9


def test_mirror_modules_expose_reference_constants_and_fail_without_cuda():
    """``mcsce_b200.libs.libsasa`` / ``libhpmf`` carry the reference modules' constants; building a calculator
    without a CUDA device raises (no CPU fallback)."""
    import torch

    from mcsce_b200.libs import libhpmf, libsasa

    assert libsasa.PIJ_BONDED == 0.8875 and libsasa.PIJ_NON_BONDED == 0.3516
    assert libsasa.R_PARAMS["CH3-SP3"] == 2.0 and libsasa.P_PARAMS["C-SP3"] == 2.149
    assert (libhpmf.C1, libhpmf.W2, libhpmf.H3, libhpmf.AC) == (3.81679, 1.39064, -0.09055, 6.0)
    with pytest.raises(AttributeError):
        libsasa.NOT_THERE
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            libsasa.calc_sasa(["C-SP3"], np.zeros((1, 1)))
        with pytest.raises(RuntimeError):
            libhpmf.init_hpmf_calculator(["C-SP3"], np.ones(1, bool), np.zeros((1, 1)))


def test_cli_flag_for_the_term():
    import inspect

    from mcsce_b200 import cli

    a = vars(cli.parser.parse_args(["x.pdb", "4", "--hpmf"]))
    assert a["hpmf"] is True and vars(cli.parser.parse_args(["x.pdb", "4"]))["hpmf"] is False
    assert set(a) == set(inspect.signature(cli.main).parameters)      # every option reaches main()
