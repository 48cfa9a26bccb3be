"""Medium-scale and edge-case GPU parity against the C oracle (fast enough to replay whole bundles):
multi-helix bundles, a multi-chain complex, retained residues (--fix), all-rigid peptides, tiny trial sets,
forced environment splits."""
import numpy as np
import pytest

from golden_utils import library

pytestmark = pytest.mark.gpu
BETA = 1.0 / (300 * 0.008314462)


def _setup(pdb_text, retain=()):
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import build_system
    from oracle.growth_c import COracle

    lib, ptm = library()
    s = Structure(pdb_text).build().remove_side_chains(list(retain))
    sysm = build_system(s, retain_idxs=tuple(retain), rotamer_library=lib, ptm_library=ptm)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    return s, sysm, eng, COracle(sysm, s.coords)


def _replay_and_compare(sysm, eng, co, n_conf, seed, env_split=0, no_graph=False):
    """GPU growth -> the chi rows it drew and the uniforms go through the C oracle -> the oracle's measured dihedrals go
    back into K1 (replay hook): coordinates bit-identical, every per-trial energy within the north-star tolerance, every
    clash flag and selection identical.  The first (production-K1) run must already agree on flags and selections."""
    rng = np.random.default_rng(seed)
    u = rng.random((n_conf, sysm.n_res))
    first = eng.grow(n_conf, BETA, seed=seed, uniforms=u, dump=True, env_split=env_split, no_graph=no_graph)
    chis = first["chis"].cpu().numpy()
    ref = co.grow(n_conf, BETA, chis=chis, uniforms=u, dump=True)
    out = eng.grow(n_conf, BETA, seed=seed, chis=chis, uniforms=u, oris=ref["ori"], dump=True, env_split=env_split,
                   no_graph=no_graph)
    grow = [st.index for st in sysm.steps if st.kind == "grow"]
    assert np.array_equal(first["selected"].cpu().numpy(), out["selected"].cpu().numpy())
    sel, ok, te = out["selected"].cpu().numpy(), out["ok"].cpu().numpy(), out["trial_energy"].cpu().numpy()
    assert np.array_equal(ok, ref["ok"])
    for c in range(n_conf):
        for i in grow:
            assert sel[c, i] == ref["selected"][c, i], (c, i, sysm.steps[i].resname)
            if sel[c, i] < 0:
                break
        o = te[c]; r = ref["trial_energy"][c]
        both = ~np.isnan(o) & ~np.isnan(r)
        assert np.array_equal(np.isinf(o[both]), np.isinf(r[both]))
        fin = both & np.isfinite(r)
        assert (np.abs(o[fin] - r[fin]) <= np.maximum(1e-3, 1e-5 * np.abs(r[fin]))).all(), (
            c, float((np.abs(o[fin] - r[fin]) / np.maximum(1e-3, 1e-5 * np.abs(r[fin]))).max()))
        if ok[c]:
            # the total accumulates one per-residue error (each within the per-trial tolerance) per grown residue
            assert abs(float(out["energy"][c]) - ref["energy"][c]) <= 4e-4 * len(grow) + 2e-5 * abs(ref["energy"][c])
            assert np.array_equal(out["coords"][c].cpu().numpy(), ref["coords"][c])
    return ok


def test_two_helix_bundle_60_residues():
    from mcsce_b200 import synth

    seq = synth.random_sequence(60, 21)
    s, sysm, eng, co = _setup(synth.helix_bundle_pdb([seq[:30], seq[30:]], seed=21, one_chain=True, spacing=12.0))
    ok = _replay_and_compare(sysm, eng, co, 6, seed=5)
    _replay_and_compare(sysm, eng, co, 3, seed=6, env_split=5)
    assert ok.shape == (6,)


def test_three_chain_complex_with_termini():
    from mcsce_b200 import synth

    seq = synth.random_sequence(45, 33, chain_starts=(0, 15, 30))
    s, sysm, eng, co = _setup(synth.helix_bundle_pdb([seq[:15], seq[15:30], seq[30:]], seed=33, one_chain=False, spacing=14.0))
    assert sysm.is_nterm.sum() > 0 and len(set(sysm.chains)) == 3
    _replay_and_compare(sysm, eng, co, 4, seed=8)


def test_retained_residues_fix_mode():
    """--fix: residues 3-5 and 9 keep the side chains of the input structure (cli.py:118-147)."""
    from mcsce_b200 import synth
    from mcsce_b200.chem.tables import load_tables
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import build_system
    from oracle.growth_c import COracle

    lib, ptm = library()
    seq = ["MET", "LYS", "PHE", "GLU", "LEU", "SER", "ARG", "ASP", "TRP", "GLN", "VAL", "ASN"]
    bb = Structure(synth.small_peptide_pdb(seq, seed=2)).build().remove_side_chains([])
    sysm0 = build_system(bb, rotamer_library=lib, ptm_library=ptm)
    eng0 = GrowthEngine(0).set_system(sysm0).set_backbone(bb.coords)
    out = eng0.grow(8, BETA, seed=4, host_out=True)
    c = int(np.argmax(out["ok"]))
    assert out["ok"][c]
    t = load_tables()
    full = bb.copy()
    for st in sysm0.steps:
        tm = t.templates[st.resname]
        full.append_atoms([tm.names[k] for k in tm.sc_idx], tm.resname, "A", st.resnum,
                          out["coords"][c][st.sc_start:st.sc_start + st.n_sc])
    full.reorder_with_resnum()
    retain = [3, 4, 5, 9]
    s = full.remove_side_chains(retain)
    assert len(s) > len(bb)
    sysm = build_system(s, retain_idxs=tuple(retain), rotamer_library=lib, ptm_library=ptm)
    assert [st.kind for st in sysm.steps if st.resnum in retain] == ["skip"] * 4
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    co = COracle(sysm, s.coords)
    _replay_and_compare(sysm, eng, co, 5, seed=12)


def test_direct_launch_regime_with_neighbour_screen_against_oracle():
    """Small systems replay a CUDA graph without the neighbour screen; with ``no_graph`` they take the path big runs
    take (screen + compaction of the surviving trials).  Same parity bar against the C oracle: a bundle, a three-chain
    complex, and a helix with retained residues (their side chains are input atoms of the screen lists)."""
    from mcsce_b200 import synth
    from mcsce_b200.chem.tables import load_tables
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import build_system
    from oracle.growth_c import COracle

    seq = synth.random_sequence(60, 21)
    s, sysm, eng, co = _setup(synth.helix_bundle_pdb([seq[:30], seq[30:]], seed=21, one_chain=True, spacing=12.0))
    eng.counters(reset=True)
    _replay_and_compare(sysm, eng, co, 6, seed=5, no_graph=True)
    _, alg, done = eng.counters(reset=True, computed=True)
    assert done < 0.85 * alg                                   # the screen really removed trials
    seq = synth.random_sequence(45, 33, chain_starts=(0, 15, 30))
    s, sysm, eng, co = _setup(synth.helix_bundle_pdb([seq[:15], seq[15:30], seq[30:]], seed=33, one_chain=False, spacing=14.0))
    _replay_and_compare(sysm, eng, co, 4, seed=8, no_graph=True)

    lib, ptm = library()
    seq = synth.random_sequence(40, 7)
    bb = Structure(synth.helix_bundle_pdb([seq], seed=7, one_chain=True)).build().remove_side_chains([])
    sysm0 = build_system(bb, rotamer_library=lib, ptm_library=ptm)
    eng0 = GrowthEngine(0).set_system(sysm0).set_backbone(bb.coords)
    out = eng0.grow(16, BETA, seed=4, host_out=True)
    c = int(np.argmax(out["ok"]))
    assert out["ok"][c]
    t = load_tables()
    full = bb.copy()
    for st in sysm0.steps:
        tm = t.templates[st.resname]
        full.append_atoms([tm.names[k] for k in tm.sc_idx], tm.resname, "A", st.resnum,
                          out["coords"][c][st.sc_start:st.sc_start + st.n_sc])
    full.reorder_with_resnum()
    retain = [5, 6, 20, 21, 33]
    s = full.remove_side_chains(retain)
    sysm = build_system(s, retain_idxs=tuple(retain), rotamer_library=lib, ptm_library=ptm)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    co = COracle(sysm, s.coords)
    _replay_and_compare(sysm, eng, co, 5, seed=12, no_graph=True)


def test_all_rigid_and_tiny_cases():
    from mcsce_b200 import synth

    # GLY/ALA only: nothing to score; every conformer is the rigid placement, energy = input-atom energy
    s, sysm, eng, co = _setup(synth.small_peptide_pdb(["GLY", "ALA", "ALA", "GLY", "ALA"], seed=1))
    out = eng.grow(3, BETA, seed=1, host_out=True)
    ref = co.grow(3, BETA)
    assert out["ok"].all() and np.allclose(out["energy"], ref["energy"], rtol=1e-12)
    assert np.abs(out["coords"] - ref["coords"]).max() <= 1e-6
    # a single grown residue between rigid ones, one conformer
    s, sysm, eng, co = _setup(synth.small_peptide_pdb(["ALA", "SER", "GLY"], seed=1))
    _replay_and_compare(sysm, eng, co, 1, seed=3)
    # many conformers of a short peptide
    s, sysm, eng, co = _setup(synth.small_peptide_pdb(["LYS", "ASP", "THR", "PHE"], seed=5))
    _replay_and_compare(sysm, eng, co, 64, seed=9)
