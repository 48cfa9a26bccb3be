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
