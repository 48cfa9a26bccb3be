"""BASELINE.json's full-size configurations on the GPU, checked through size-independent properties:
the energy the fused growth accumulated equals the C oracle's level-by-level re-scoring of the final
coordinates, no scored pair of a successful conformer is inside its clash distance, runs are
deterministic, results do not depend on how conformers are sharded, and a replay of whole conformers
through the C oracle reproduces every selection."""
import numpy as np
import pytest

from golden_utils import library

pytestmark = pytest.mark.gpu
BETA = 1.0 / (300 * 0.008314462)
_CACHE = {}


def _workload(cfg):
    if cfg in _CACHE:
        return _CACHE[cfg]
    from mcsce_b200 import synth
    from mcsce_b200.engine import GrowthEngine
    from mcsce_b200.structure import Structure
    from mcsce_b200.system import build_system
    from oracle.growth_c import COracle

    lib, ptm = library()
    s = Structure(synth.config_backbone(cfg)).build().remove_side_chains([])
    sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    _CACHE[cfg] = (s, sysm, eng, COracle(sysm, s.coords))
    return _CACHE[cfg]


def _check_rescoring(sysm, co, coords, energy, ok, sample, n_grow):
    checked = 0
    for c in sample:
        if not ok[c]:
            continue
        e, ratio = co.rescore(coords[c])
        assert abs(energy[c] - e) <= 4e-4 * n_grow + 2e-5 * abs(e), (c, energy[c], e)
        assert ratio >= 1.0 - 1e-6, (c, ratio)
        checked += 1
    return checked


def test_config3_300_residues_2048_conformers():
    s, sysm, eng, co = _workload(3)
    n_grow = sum(st.kind == "grow" for st in sysm.steps)
    a = eng.grow(2048, BETA, seed=42, host_out=True)
    coords, energy, ok = a["coords"].copy(), a["energy"].copy(), a["ok"].copy()
    assert ok.mean() > 0.8 and np.isfinite(energy[ok == 1]).all() and np.isnan(energy[ok == 0]).all()
    assert np.isfinite(coords[ok == 1]).all()
    # input atoms untouched, every grown conformer different
    assert np.array_equal(coords[:, :sysm.n_input], np.broadcast_to(s.coords, (2048,) + s.coords.shape))
    assert len(np.unique(np.round(energy[ok == 1], 6))) > 0.99 * ok.sum()
    rng = np.random.default_rng(0)
    assert _check_rescoring(sysm, co, coords, energy, ok, rng.choice(2048, 10, replace=False), n_grow) >= 6
    # determinism and sharding invariance (conformer streams are keyed by the global id)
    b = eng.grow(2048, BETA, seed=42, host_out=True)
    assert np.array_equal(b["energy"], energy, equal_nan=True) and np.array_equal(b["coords"], coords)
    lo = eng.grow(1024, BETA, seed=42, conf_id0=0, host_out=True)
    e_lo, c_lo = lo["energy"].copy(), lo["coords"].copy()
    hi = eng.grow(1024, BETA, seed=42, conf_id0=1024, host_out=True)
    assert np.array_equal(np.concatenate([e_lo, hi["energy"]]), energy, equal_nan=True)
    assert np.array_equal(np.concatenate([c_lo, hi["coords"]]), coords)
    c = eng.grow(64, BETA, seed=43, host_out=True)
    assert not np.array_equal(c["energy"], energy[:64], equal_nan=True)


def _replay_whole_conformers(cfg, n, seed, min_selections):
    """Grow n conformers on the GPU, feed the chi rows it drew and the same uniforms to the C oracle: every selection
    and clash decision must agree.  Per-trial energies are compared twice: (1) with the C oracle's measured dihedrals fed
    back to K1 (the replay hook that pins K1: coordinates then agree bit for bit) EVERY trial must be within the north-star
    tolerance; (2) the production K1 (template-constant dihedrals) against the oracle's float32 re-measurement: measured
    bounds on the fraction outside and on the worst trial (DESIGN "K1 fidelity")."""
    s, sysm, eng, co = _workload(cfg)
    rng = np.random.default_rng(5)
    u = rng.random((n, sysm.n_res))
    out = eng.grow(n, BETA, seed=seed, uniforms=u, dump=True)
    chis = out["chis"].cpu().numpy()
    ref = co.grow(n, BETA, chis=chis, uniforms=u, dump=True)
    pinned = eng.grow(n, BETA, seed=seed, chis=chis, uniforms=u, oris=ref["ori"], dump=True)
    grow = [st.index for st in sysm.steps if st.kind == "grow"]
    stats = {}
    for tag, o in (("constant", out), ("pinned", pinned)):
        sel = o["selected"].cpu().numpy()
        assert np.array_equal(o["ok"].cpu().numpy(), ref["ok"])
        n_sel = 0
        for c in range(n):
            for i in grow:
                assert sel[c, i] == ref["selected"][c, i], (tag, c, i, sysm.steps[i].resname)
                if sel[c, i] < 0:
                    break
                n_sel += 1
        assert n_sel > min_selections
        te, tr = o["trial_energy"].cpu().numpy(), ref["trial_energy"]
        both = ~np.isnan(te) & ~np.isnan(tr)
        assert np.array_equal(np.isinf(te[both]), np.isinf(tr[both]))
        fin = both & np.isfinite(tr)
        ratio = np.abs(te[fin] - tr[fin]) / np.maximum(1e-3, 1e-5 * np.abs(tr[fin]))
        stats[tag] = (int(fin.sum()), int((ratio > 1).sum()), float(ratio.max()))
        print("cfg%s replay (%s dihedrals): %d trials, %d selections, %d trials outside tolerance, worst %.2fx" % (
            cfg, tag, fin.sum(), n_sel, (ratio > 1).sum(), ratio.max()))
        if ref["ok"].any():
            c = int(np.where(ref["ok"])[0][0])
            dc = np.abs(o["coords"][c].cpu().numpy() - ref["coords"][c])
            if tag == "pinned":
                assert dc.max() == 0.0, "K1 fed the oracle's measured dihedrals must reproduce its coordinates bit for bit"
    assert stats["pinned"][1] == 0 and stats["pinned"][2] <= 1.0, stats      # measured worst: 0.60x (config 3), 0.78x (config 4)
    # production K1: bounds measured on B200 (round 2) - fraction outside <= 2 %, none beyond 20x (measured worst 9.4x-11.1x, an ARG rotamer with a steep contact inside the side chain)
    assert stats["constant"][1] <= 0.02 * stats["constant"][0] and stats["constant"][2] < 20.0, stats


@pytest.mark.parametrize("cfg,n", [(3, 8), (4, 16)])
def test_scored_energy_of_every_selected_trial_matches_oracle_on_the_same_coordinates(cfg, n):
    """K2 + K2s at full size with no K1 in the comparison: the energy the growth scored for the trial it selected, per
    residue, against the C oracle's float64 re-scoring of that residue on the GPU's own final coordinates.  North-star
    tolerance per trial: max(1e-3 kJ/mol, 1e-5 |E|); measured worst 0.31x (config 3), the float32 partial sums between
    two float64 flushes being the main term (FLUSH = 32 pairs; 64 doubles it, 128 reaches the tolerance)."""
    s, sysm, eng, co = _workload(cfg)
    out = eng.grow(n, BETA, seed=7, dump=True, no_graph=True)
    te, sel, ok, xyz = (out[k].cpu().numpy() for k in ("trial_energy", "selected", "ok", "coords"))
    ratios = []
    for c in np.where(ok)[0]:
        total, clash_ratio, lev = co.rescore_levels(xyz[c])
        assert clash_ratio >= 1.0 - 1e-6
        for st in sysm.steps:
            if st.kind != "grow":
                continue
            e_gpu, e_ref = te[c, eng.trial_off[st.index] + sel[c, st.index]], lev[st.index]
            ratios.append(abs(e_gpu - e_ref) / max(1e-3, 1e-5 * abs(e_ref)))
    ratios = np.array(ratios)
    print("cfg%d: %d residues, |dE|/tolerance max %.3f p99 %.3f median %.4f" % (cfg, len(ratios), ratios.max(),
                                                                                  np.quantile(ratios, 0.99), np.median(ratios)))
    assert len(ratios) > 500 and ratios.max() < 0.6


def _independent_energy_functions(s, picks, tables):
    """{i: (ref_numpy.EnergyFunction of residue ordinal i's side chain against everything placed before it, n_old)},
    built from the Structure's own atom bookkeeping (labels, residue numbers, termini) the way the reference's
    initialize_func_calc walks the levels (side_chain_builder.py:42-107) - nothing from mcsce_b200.system."""
    from oracle import ref_numpy as O

    work = s.copy()
    first = int(work.res_nums[0])
    types, chains = work.residue_types, work.residue_chains
    out = {}
    for j in range(max(picks) + 1):
        tm = tables.templates[types[j]]
        work.append_atoms([tm.names[k] for k in tm.sc_idx], tm.resname, chains[j], first + j, np.zeros((tm.n_sc, 3), np.float32))
        if j in picks:
            n_all = len(work)
            n_term, c_term = work.get_terminal_res_atom_arr()
            out[j] = (O.EnergyFunction(work.atom_labels, work.res_nums, work.res_labels, n_term, c_term,
                                       np.arange(n_all - tm.n_sc, n_all), np.arange(n_all - tm.n_sc), tables=tables),
                      n_all - tm.n_sc)
    return out


@pytest.mark.parametrize("cfg,n,n_pick", [(3, 4, 12), (4, 12, 14), (5, 4, 12)])
def test_per_trial_energies_against_the_table_independent_numpy_oracle(cfg, n, n_pick):
    """Full-size check that shares NO host table with the product: for chain termini, first/last residues of chains, PTM
    residues and random interior residues, ``oracle.ref_numpy.EnergyFunction`` (per-pair coefficient arrays from the
    reference's rules, libenergy.py:27-233) scores the GPU's own trial coordinates against the GPU's environment;
    every per-trial energy must be within the north-star tolerance of what the fused growth recorded, every clash flag equal."""
    from mcsce_b200.chem.tables import load_tables

    s, sysm, eng, co = _workload(cfg)
    t = load_tables()
    out = eng.grow(n, BETA, seed=31, dump=True)
    te, chis, ok, xyz = (out[k].cpu().numpy() for k in ("trial_energy", "chis", "ok", "coords"))
    assert ok.any()
    c = int(np.where(ok)[0][0])
    grown = [st for st in sysm.steps if st.kind == "grow"]
    chains = np.asarray(sysm.chains)
    res_chain = [chains[st.sc_start] for st in grown]
    pick = set()
    for k, st in enumerate(grown):          # first / last grown residue of every chain
        if k == 0 or k == len(grown) - 1 or res_chain[k] != res_chain[k - 1] or res_chain[k] != res_chain[k + 1]:
            pick.add(st.index)
    ptm_names = ("SEP", "TPO", "PTR", "H2D", "H2E", "M3L", "KCX", "HYP")
    seen = set()
    for st in grown:                         # one of each PTM residue type
        if st.resname in ptm_names and st.resname not in seen:
            seen.add(st.resname); pick.add(st.index)
    rng = np.random.default_rng(cfg)
    rest = [st.index for st in grown if st.index not in pick]
    pick = sorted(pick)[:max(0, n_pick - 4)]
    pick = pick + sorted(rng.choice(rest, n_pick - len(pick), replace=False).tolist())
    worst = 0.0; n_trials = 0
    efs = _independent_energy_functions(s, set(pick), t)
    for i in sorted(set(pick)):
        st = sysm.steps[i]
        ef, n_old = efs[i]
        assert n_old == st.sc_start
        o, T = int(eng.trial_off[i]), len(st.prob)
        trial = eng.place_trials(i, chis[c, o:o + T]).cpu().numpy()
        env = xyz[c, :n_old]
        ref = ef.calculate(trial, np.broadcast_to(env[None], (T,) + env.shape))
        got = te[c, o:o + T]
        assert np.array_equal(np.isinf(got), np.isinf(ref)), (cfg, i, st.resname)
        fin = np.isfinite(ref)
        if fin.any():
            r = np.abs(got[fin] - ref[fin]) / np.maximum(1e-3, 1e-5 * np.abs(ref[fin]))
            worst = max(worst, float(r.max())); n_trials += int(fin.sum())
            assert r.max() <= 1.0, (cfg, i, st.resname, float(r.max()))
    print("cfg%d vs table-independent oracle: %d residues, %d finite trials, worst %.3fx tolerance" % (cfg, len(set(pick)), n_trials, worst))
    assert n_trials > 100


def test_conformer_groups_in_flight_keep_per_conformer_buffers_apart():
    """Calls of >= 192 conformers grow two groups side by side on separate streams; dumps and replay inputs are
    per-conformer arrays that each group must address with its own offset.  (The environment split is pinned to 1: the
    automatic rule depends on the number of conformers in the call, and a different split regroups float32 partial sums;
    whole and halves are both >= 1024 conformers, so both run the per-conformer scoring pass - the per-tile kernel of smaller
    calls groups the sums differently.)"""
    s, sysm, eng, co = _workload(3)
    n, half = 2304, 1152
    u = np.random.default_rng(3).random((n, sysm.n_res))
    whole = eng.grow(n, BETA, seed=17, uniforms=u, dump=True, env_split=1)
    sel, te, ch = (whole[k].cpu().numpy().copy() for k in ("selected", "trial_energy", "chis"))
    en, ok = whole["energy"].cpu().numpy().copy(), whole["ok"].cpu().numpy().copy()
    del whole
    for lo in (0, half):
        part = eng.grow(half, BETA, seed=17, conf_id0=lo, uniforms=u[lo:lo + half], dump=True, env_split=1)
        assert np.array_equal(part["selected"].cpu().numpy(), sel[lo:lo + half])
        assert np.array_equal(part["trial_energy"].cpu().numpy(), te[lo:lo + half], equal_nan=True)
        # (K1 of residue i+1 runs ahead of the selection of residue i on the second stream: whether a conformer that dies at
        #  residue i still gets its rows of residue i+1 written is a matter of timing - compare the rows both runs wrote)
        pc, wc = part["chis"].cpu().numpy(), ch[lo:lo + half]
        both = (pc != 0) & (wc != 0)
        assert np.array_equal(pc[both], wc[both]) and both.sum() > 0.9 * (wc != 0).sum()
        assert np.array_equal(part["energy"].cpu().numpy(), en[lo:lo + half], equal_nan=True)
        assert np.array_equal(part["ok"].cpu().numpy(), ok[lo:lo + half])
        del part
    # replaying the recorded chi rows through the explicit-input path gives the same growth
    again = eng.grow(n, BETA, seed=99, uniforms=u, chis=ch, dump=True, env_split=1)
    assert np.array_equal(again["selected"].cpu().numpy(), sel)


def test_config3_whole_conformer_replay_through_c_oracle():
    _replay_whole_conformers(3, 6, 11, 1000)


def test_config4_150_residues_with_ptm_residues():
    """SEP, TPO, PTR, H2D, H2E, M3L, KCX, HYP (three of each) among 150 residues: extended rotamer / parameter tables."""
    s, sysm, eng, co = _workload(4)
    names = [st.resname for st in sysm.steps]
    for r in ("SEP", "TPO", "PTR", "H2D", "H2E", "M3L", "KCX", "HYP"):
        assert names.count(r) >= 3, r
    n_grow = sum(st.kind == "grow" for st in sysm.steps)
    a = eng.grow(512, BETA, seed=21, host_out=True)
    coords, energy, ok = a["coords"].copy(), a["energy"].copy(), a["ok"].copy()
    assert ok.mean() > 0.2 and np.isfinite(energy[ok == 1]).all()
    assert _check_rescoring(sysm, co, coords, energy, ok, list(np.where(ok)[0][:8]), n_grow) == 8
    b = eng.grow(512, BETA, seed=21, host_out=True)
    assert np.array_equal(b["energy"], energy, equal_nan=True) and np.array_equal(b["coords"], coords)
    _replay_whole_conformers(4, 8, 13, 400)


def test_config5_2000_residues_8_chains():
    s, sysm, eng, co = _workload(5)
    assert sysm.n_res == 2000 and len(set(sysm.chains)) == 8
    n_grow = sum(st.kind == "grow" for st in sysm.steps)
    a = eng.grow(16, BETA, seed=3, host_out=True)
    coords, energy, ok = a["coords"].copy(), a["energy"].copy(), a["ok"].copy()
    assert ok.sum() >= 2
    assert _check_rescoring(sysm, co, coords, energy, ok, list(np.where(ok)[0][:2]), n_grow) == 2
    # A different environment split regroups the float32 partial sums (1e-7 relative per trial energy), so the
    # conformers agree to rounding; over 1800 residues x 16 conformers a selection whose uniform sits within that
    # distance of a boundary may flip and send one conformer down another path.
    b = eng.grow(16, BETA, seed=3, env_split=7, host_out=True)
    same = np.isclose(b["energy"], energy, rtol=2e-6, atol=0, equal_nan=True)
    assert same.mean() >= 0.8, (b["energy"], energy)
    # the same split twice is bitwise reproducible
    b2 = eng.grow(16, BETA, seed=3, env_split=7, host_out=True)
    assert np.array_equal(b2["energy"], b["energy"], equal_nan=True)


def test_config2_76_residues_128_conformers():
    s, sysm, eng, co = _workload(2)
    n_grow = sum(st.kind == "grow" for st in sysm.steps)
    a = eng.grow(128, BETA, seed=9, host_out=True)
    assert a["ok"].mean() > 0.9
    assert _check_rescoring(sysm, co, a["coords"], a["energy"], a["ok"], range(0, 128, 8), n_grow) >= 12
