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
