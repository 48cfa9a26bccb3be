"""GPU tests of the reference-facing drop-in surface: the Python entry points keep the
reference's signatures and return conventions while the work runs in the CUDA library."""
import os

import numpy as np
import pytest

from golden_utils import GOLDEN, GoldenCase, library

pytestmark = pytest.mark.gpu


def test_rotate_sidechain_matches_reference_fixture():
    from mcsce_b200.chem.tables import load_tables
    from mcsce_b200.libs.libscbuild import rotate_sidechain, rotate_sidechain_batch

    z = np.load(GOLDEN / "templates.npz")
    t = load_tables()
    n_exact = n_tot = 0
    for res, tm in t.templates.items():
        if res in ("ALA", "GLY"):
            with pytest.raises(AssertionError):
                rotate_sidechain(res, np.array([10.0]))
            continue
        for n in range(1, 7):
            key = "%s_n%d_tors" % (res, n)
            if key not in z.files or n > tm.n_sound:
                break
            got, sc_idx = rotate_sidechain_batch(res, z[key])
            ref = z["%s_n%d_rot" % (res, n)]
            assert np.array_equal(sc_idx, tm.sc_idx)
            assert got.dtype == np.float32 and got.shape == ref.shape
            assert np.abs(got - ref).max() <= 4e-6, (res, n, np.abs(got - ref).max())
            if n == 1:                    # one rotation: no float32 re-measurement noise, expected bit-exact
                n_exact += int((got == ref).all()); n_tot += 1
    single, _ = rotate_sidechain("LYS", z["LYS_n4_tors"][0])
    assert np.abs(single - z["LYS_n4_rot"][0]).max() <= 4e-6
    assert n_exact >= 0.9 * n_tot, (n_exact, n_tot)


def test_place_sidechain_template_matches_reference_fixture():
    from mcsce_b200.chem.tables import load_tables
    from mcsce_b200.libs.libcalc import place_sidechain_template

    z = np.load(GOLDEN / "templates.npz")
    t = load_tables()
    for res, tm in t.templates.items():
        got = place_sidechain_template(z[res + "_bb"], tm.coords)
        assert got.dtype == np.float64 and got.shape == (tm.n_atoms, 3)
        assert np.abs(got - z[res + "_placed"]).max() <= 1e-9, res


@pytest.mark.parametrize("name", ["pep6", "pep6_ljclash", "twochain"])
def test_prepare_energy_function_closure(name):
    """The closure contract of libenergy.py:786-810 on the reference's own trial coordinates."""
    from test_gpu_parity import oracle_env_sequence, within_tol
    from mcsce_b200.chem.tables import load_tables
    from mcsce_b200.libs.libenergy import prepare_energy_function
    from mcsce_b200.system import build_system

    case = GoldenCase(name)
    lib, ptm = library()
    s = case.structure()
    sysm = build_system(s, terms=case.terms, rotamer_library=lib, ptm_library=ptm)
    t = load_tables()
    envs = oracle_env_sequence(case, 0, s, sysm)
    steps = case.steps_by_level(0)
    # level 0: input atoms alone
    n_term, c_term = s.get_terminal_res_atom_arr()
    f0 = prepare_energy_function(s.atom_labels, s.res_nums, s.res_labels, n_term, c_term, None, terms=list(case.terms))
    e0 = f0(s.coords[None], s.coords[None, :0])
    assert e0.shape == (1,) and abs(e0[0] - float(case.conf(0)["e_backbone"])) <= 1e-9 * max(1, abs(e0[0]))
    work = s.copy()
    checked = 0
    for st in sysm.steps:
        tm = t.templates[st.resname]
        work.append_atoms([tm.names[k] for k in tm.sc_idx], tm.resname, str(sysm.chains[st.sc_start]), st.resnum, np.zeros((tm.n_sc, 3)))
        if st.index not in envs or checked >= 3:
            continue
        n_term, c_term = work.get_terminal_res_atom_arr()
        idx = np.arange(len(work))
        f = prepare_energy_function(work.atom_labels, work.res_nums, work.res_labels, n_term, c_term, None, batch_size=16,
                                    partial_indices=[idx[-st.n_sc:], idx[:-st.n_sc]], terms=list(case.terms))
        g = steps[st.index]
        trial = g["trial"]
        env = np.repeat(envs[st.index][None], len(trial), axis=0)
        got = f(trial, env)                              # identical rows -> one environment, T trials
        assert got.dtype == np.float64 and got.shape == (len(trial),)
        assert np.array_equal(np.isinf(got), np.isinf(g["energies"]))
        fin = ~np.isinf(g["energies"])
        assert within_tol(got[fin], g["energies"][fin]).all()
        env2 = env.copy(); env2[0, 0, 0] += 1e-3        # rows differ -> one environment per trial
        got2 = f(trial[:8], env2[:8])
        assert np.array_equal(np.isinf(got2[1:]), np.isinf(g["energies"][1:8]))
        checked += 1
    with pytest.raises(NotImplementedError):
        prepare_energy_function(s.atom_labels, s.res_nums, s.res_labels, n_term[:len(s)], c_term[:len(s)], None, terms=["lj", "gb"])
    with pytest.raises(AssertionError):
        f0(s.coords, s.coords)


def test_total_energy_of_grown_conformers_matches_oracle_rescoring():
    """Independent of any replay: the accumulated energy the fused growth reports must equal the oracle's
    energy functions evaluated level by level on the final coordinates."""
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.system import build_system
    from oracle import ref_numpy as O

    case = GoldenCase("pep12")
    lib, ptm = library()
    s = case.structure()
    sysm = build_system(s, terms=case.terms, rotamer_library=lib, ptm_library=ptm)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    out = eng.grow(16, case.beta, seed=21, host_out=True)
    efs = O.prepare_energy_functions(s, terms=case.terms)
    n_checked = 0
    for c in range(16):
        if not out["ok"][c]:
            continue
        xyz = out["coords"][c]
        total = efs[0].calculate(xyz[None, :sysm.n_input], xyz[None, :0])[0]
        for st in sysm.steps:
            if st.kind == "grow":
                new = xyz[st.sc_start:st.sc_start + st.n_sc]
                total += efs[st.index + 1].calculate(new[None], xyz[None, :st.sc_start])[0]
        assert abs(out["energy"][c] - total) <= 1e-3 + 1e-5 * abs(total), (c, out["energy"][c], total)
        n_checked += 1
    assert n_checked >= 8


def test_side_chain_builder_entry_points(tmp_path):
    from functools import partial

    from mcsce_b200.core import side_chain_builder as scb
    from mcsce_b200.libs.libenergy import prepare_energy_function
    from mcsce_b200.structure import Structure

    case = GoldenCase("pep12")
    s = case.structure()
    scb.set_seed(5)
    scb.initialize_func_calc(partial(prepare_energy_function, batch_size=16, forcefield=None, terms=["lj", "clash", "coulomb"]), structure=s)
    assert len(scb.sidechain_placeholders) == len(s.residue_types) + 1 == len(scb.energy_calculators)
    beta = case.beta
    st, ok, energy, addr = scb.create_side_chain_structure([s.coords, beta, [], str(tmp_path / "one.pdb")])
    assert ok and st is not None and np.isfinite(energy) and os.path.exists(addr)
    assert len(st) == len(scb.sidechain_placeholders[-1])
    assert list(st.res_nums) == sorted(st.res_nums)
    confs, success = scb.create_side_chain_ensemble(s, 12, 300, str(tmp_path), parallel_worker=4)
    assert len(confs) == len(success) == 12 and sum(success) >= 6
    lines = (tmp_path / "energies.csv").read_text().strip().splitlines()
    assert lines[0] == "File name,Energy(kJ/mol)" and len(lines) == 1 + sum(success)
    for n, good in enumerate(success):
        assert os.path.exists(tmp_path / ("%d.pdb" % n)) == bool(good)
    re_read = Structure(str(tmp_path / ("%d.pdb" % success.index(True)))).build()
    assert len(re_read) == len(st)
    best = scb.create_side_chain(s, 8, 300, return_first_valid=False)
    first = scb.create_side_chain(s, 8, 300, return_first_valid=True)
    assert best is not None and first is not None and len(best) == len(st)
    # the mode reductions run on the device (mcsce_pick): same picks as the host reductions of the reference
    # (side_chain_builder.py:245-249 first valid, :280-281 lowest energy), only the picked conformer is copied back
    eng = scb._state["engine"]
    out = eng.grow(64, beta, seed=3)
    e, k, x = out["energy"].cpu().numpy(), out["ok"].cpu().numpy(), out["coords"].cpu().numpy()
    k[::3] = 0                                         # (some failures, so that the first valid one is not conformer 0)
    out["ok"].copy_(out["ok"].new_tensor(k))
    i_best, x_best = eng.pick(out)
    i_first, x_first = eng.pick(out, first_valid=True)
    assert i_best == int(np.where(k == 1, e, np.inf).argmin()) and np.array_equal(x_best, x[i_best])
    assert i_first == int(np.argmax(k == 1)) == 1 and np.array_equal(x_first, x[1])
    out["ok"].zero_()
    assert eng.pick(out) == (-1, None)
    # a level's energy closure, built on demand, agrees with the growth's own scoring conventions
    f = scb.energy_calculators[1]
    assert scb.energy_calculators[3] is None or callable(scb.energy_calculators[3])   # GLY/ALA levels are None
    assert callable(f)


def test_one_preparation_serves_other_backbone_conformations():
    """``create_side_chain_structure`` takes the backbone per call (side_chain_builder.py:143-147): growing conformation B
    after preparing the sequence on conformation A equals preparing on B."""
    from functools import partial

    from mcsce_b200 import synth
    from mcsce_b200.core import side_chain_builder as scb
    from mcsce_b200.libs.libenergy import prepare_energy_function
    from mcsce_b200.structure import Structure

    seq = ["MET", "LYS", "PHE", "GLU", "LEU", "SER", "ARG", "ASP", "TRP", "GLN", "VAL", "ASN"]
    a = Structure(synth.small_peptide_pdb(seq, seed=2)).build().remove_side_chains([])
    b = Structure(synth.small_peptide_pdb(seq, seed=2, phi=-120.0, psi=130.0)).build().remove_side_chains([])
    creator = partial(prepare_energy_function, batch_size=16, forcefield=None, terms=["lj", "clash", "coulomb"])
    beta = 1.0 / (300 * 0.008314462)

    def grow_b(prepare_on):
        scb.initialize_func_calc(creator, structure=prepare_on)
        scb.set_seed(11)
        out = [scb.create_side_chain_structure([b.coords, beta, [], None]) for _ in range(4)]
        return [(ok, e, None if st is None else st.coords.copy()) for st, ok, e, _ in out]

    fresh, reused = grow_b(b), grow_b(a)
    assert any(ok for ok, _, _ in fresh)
    for (ok1, e1, x1), (ok2, e2, x2) in zip(fresh, reused):
        assert ok1 == ok2
        if ok1:
            assert e1 == e2 and np.array_equal(x1, x2)
            assert np.abs(x1[:4] - b.coords[:4]).max() == 0          # grown on B's backbone


def test_cli_and_bridge(tmp_path):
    from mcsce_b200 import cli
    from mcsce_b200.mcsce_sidechain import mcsce_sidechain

    case = GoldenCase("pep6")
    pdb = tmp_path / "pep6.pdb"
    pdb.write_text(case.pdb)
    cli.main(str(pdb), 6, None, str(tmp_path), str(tmp_path / "log.csv"), "ensemble", None, batch_size=16, seed=3)
    out_dir = tmp_path / "pep6_mcsce"
    log = (tmp_path / "log.csv").read_text().strip().splitlines()
    assert log[0] == "PDB name,Succeeded,Time used" and len(log) == 2
    n_ok = int(log[1].split(",")[1])
    assert n_ok == len(list(out_dir.glob("*.pdb"))) and (out_dir / "energies.csv").exists()
    cli.main(str(pdb), 6, None, str(tmp_path), str(tmp_path / "log2.csv"), "exhaustive", None, batch_size=16, seed=3)
    assert (tmp_path / "pep6_mcsce.pdb").exists()
    s = case.structure()
    xyz = mcsce_sidechain("MQCKMD", s.coords, n_trials=8, mode="exhaustive")
    assert xyz is not None and xyz.shape[1] == 3 and xyz.shape[0] > len(s)
    with pytest.raises(RuntimeError):
        mcsce_sidechain("MQCKMD", s.coords, mode="nope")


def test_large_ensembles_are_grown_in_chunks():
    from functools import partial

    from mcsce_b200.core import side_chain_builder as scb
    from mcsce_b200.libs.libenergy import prepare_energy_function

    case = GoldenCase("pep6")
    s = case.structure()
    scb.initialize_func_calc(partial(prepare_energy_function, batch_size=16, forcefield=None, terms=["lj", "clash", "coulomb"]), structure=s)
    scb.set_seed(11)
    whole = scb._grow(s.coords, case.beta, 13)
    scb.set_seed(11)
    old = scb.MAX_CONFORMERS_PER_LAUNCH
    scb.MAX_CONFORMERS_PER_LAUNCH = 5
    try:
        chunked = scb._grow(s.coords, case.beta, 13)
    finally:
        scb.MAX_CONFORMERS_PER_LAUNCH = old
    # the same Philox streams (keyed by global conformer id) whatever the chunking: same selections, hence bit-identical
    # coordinates; the energies agree to the float32 grouping of the pair sums, which follows the launch geometry (13
    # conformers get clusters of 8 CTAs each, 5 get clusters of 16; in the per-kernel path it follows the environment split)
    for a, b in zip(whole, chunked):
        a, b = np.asarray(a), np.asarray(b)
        if a.dtype == np.float64:
            assert np.allclose(a, b, rtol=1e-6, atol=1e-3, equal_nan=True)
        else:
            assert np.array_equal(a, b, equal_nan=True)
