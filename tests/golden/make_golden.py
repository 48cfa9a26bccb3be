"""Generate the golden fixtures by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py [case ...]

The reference has no tests or known-answer vectors of its own, so these files are
the pin for the oracle (``oracle/ref_numpy.py``) and, through it, for the CUDA path.
Nothing in the reference is edited: its module-level callables are *wrapped* with
recorders (``rotate_sidechain`` -> perturbed chi rows, ``libscbuild.calc_torsion_angles`` -> the
float32 dihedral it re-measures before every chi rotation, the per-residue energy closures
-> trial coordinates and energies, ``choose_random`` -> Boltzmann weights and picks).
The uniforms that ``choose_random`` consumes come from numba's own generator
(``side_chain_builder.py:36``); they are recovered by seeding that generator from a
jitted helper, drawing the sequence, and re-seeding before the run.  The in-place
``+=`` on the library's cached rows (quirk Q1, ``side_chain_builder.py:173``) is undone
after every conformer so each run starts from the pristine library.

Inputs are produced by ``mcsce_b200.synth`` (seeded), so a fixture stores the PDB text
it was made from and can be regenerated bit-for-bit.
"""
import copy
import io
import os
import sys
import tempfile
import time
from functools import partial
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))

from ref_loader import load_reference  # noqa: E402

CASES = {
    "pep6": dict(seq=["MET", "GLN", "CYS", "LYS", "HIP", "ASP"], n_conf=2, seed=11, full=True),
    "pep6_ljclash": dict(seq=["MET", "GLN", "CYS", "LYS", "HIP", "ASP"], n_conf=1, seed=12, full=True,
                         terms=["lj", "clash"]),
    "pep12": dict(seq=["MET", "ARG", "GLY", "TRP", "PRO", "ALA", "LYS", "HIP", "ASP", "PHE", "SER", "ILE"],
                  n_conf=2, seed=13, full=True),
    "ptm14": dict(seq=["SER", "SEP", "LEU", "TPO", "GLY", "PTR", "ALA", "H2D", "VAL", "H2E", "M3L", "KCX",
                       "HYP", "GLU"], n_conf=1, seed=14, full=True),
    "twochain": dict(seq=None, n_conf=1, seed=15, full=True),
    "std20": dict(seq=["THR", "ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HID", "HIE", "ILE",
                       "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL", "HIP", "ASN"],
                  n_conf=1, seed=16, full=False),
    "helix76": dict(cfg=1, n_conf=1, seed=17, full=False),
    # the eleven other PTM residues the reference can grow: S1P H1D H1E AGM ALY MEN BHD SMC MGN CGU LYZ
    "ptm18": dict(seq=["SER", "S1P", "ALA", "H1D", "GLY", "H1E", "AGM", "LEU", "ALY", "MEN", "VAL", "BHD", "SMC", "ALA",
                       "MGN", "CGU", "LYZ", "GLU"], n_conf=2, seed=19, full=True),
    # two 20-residue helices 11 A apart in one chain: trials meet atoms of the other helix (clashes that are not local in sequence)
    "bundle40": dict(seq=None, n_conf=1, seed=20, full=False),
    # BASELINE configs 4 and 3 at full size, one conformer each, light records (no trial coordinates): the stock reference's
    # per-trial energies / selections at 150 residues with PTMs and at 300 residues (init: minutes to an hour on one core)
    "cfg4_150": dict(cfg=4, n_conf=1, seed=21, full=False, skip_failed=True),
    "cfg3_300": dict(cfg=3, n_conf=1, seed=22, full=False, skip_failed=True),
    # --fix mode (cli.py:118-147): residues 3, 4 and 9 keep the side chains of an earlier growth of pep12
    "fix12": dict(base="pep12", retain=[3, 4, 9], n_conf=2, seed=18, full=True),
}


def case_pdb(name, spec):
    from mcsce_b200 import synth

    if spec.get("cfg"):
        return synth.config_backbone(spec["cfg"])
    if name == "bundle40":
        seq = synth.random_sequence(40, 20)
        return synth.helix_bundle_pdb([seq[:20], seq[20:]], seed=20, one_chain=True, spacing=11.0)
    if name == "twochain":
        a = ["LYS", "GLU", "PHE", "SER", "TRP", "ASN"]
        b = ["ARG", "LEU", "ASP", "TYR", "MET", "VAL"]
        return synth.helix_bundle_pdb([a, b], seed=15, one_chain=False, spacing=12.0)
    return synth.small_peptide_pdb(spec["seq"], seed=spec["seed"])


def run_case(name, spec, out_dir):
    import numba

    mcsce = load_reference(0)
    from mcsce.core import side_chain_builder as scb
    from mcsce.core.build_definitions import forcefields
    from mcsce.libs.libenergy import prepare_energy_function
    from mcsce.libs.libstructure import Structure

    @numba.njit
    def nb_seed(s):
        np.random.seed(s)

    @numba.njit
    def nb_draw(n):
        out = np.empty(n)
        for i in range(n):
            out[i] = np.random.random()
        return out

    terms = spec.get("terms", ["lj", "clash", "coulomb"])
    retain = list(spec.get("retain", []))
    tmp = tempfile.mkdtemp()
    pdb_path = os.path.join(tmp, name + ".pdb")
    ff = forcefields["Amberff14SB"](Cterminal="OXT", Nterminal="HN")
    if spec.get("base"):
        # input of a --fix run: a complete structure, grown here by the reference from the base case's backbone
        base_path = os.path.join(tmp, "base.pdb")
        Path(base_path).write_text(case_pdb(spec["base"], CASES[spec["base"]]))
        b = Structure(base_path)
        b.build()
        b = b.remove_side_chains([])
        scb.initialize_func_calc(partial(prepare_energy_function, batch_size=16, forcefield=ff, terms=terms), structure=b)
        np.random.seed(spec["seed"]); nb_seed(spec["seed"])
        pristine0 = copy.deepcopy(scb._rotamer_library._data)
        full_structure = None
        for _ in range(20):
            full_structure, ok, _, _ = scb.create_side_chain_structure([b.coords, 1 / (300 * 0.008314462), [], None])
            if ok:
                break
        scb._rotamer_library._data = pristine0
        assert full_structure is not None
        full_structure.write_PDB(pdb_path)
        pdb_text = Path(pdb_path).read_text()
    else:
        pdb_text = case_pdb(name, spec)
        Path(pdb_path).write_text(pdb_text)
    s = Structure(pdb_path)
    s.build()
    assert s.check_backbone_atom_completeness() == []
    s = s.remove_side_chains(retain)
    t0 = time.time()
    scb.initialize_func_calc(partial(prepare_energy_function, batch_size=16, forcefield=ff, terms=terms), structure=s,
                             retain_idxs=retain)
    print("[%s] reference init %.1fs" % (name, time.time() - t0))

    # ---- recorders (wrap, never edit) ----
    rec = {}
    orig_rotate = scb.rotate_sidechain
    orig_choose = scb.choose_random
    orig_efuncs = list(scb.energy_calculators)

    import mcsce.libs.libscbuild as lsb
    orig_torsion = lsb.calc_torsion_angles
    measured = []

    def torsion_rec(coords):
        # the float32 dihedral rotate_sidechain re-measures before every chi rotation (libscbuild.py:113,167,183,197)
        v = orig_torsion(coords)
        measured.append(float(v[0]))
        return v

    def rotate_rec(res_type, tors):
        rec["chis"].append(np.array(tors, dtype=np.float64))
        measured.clear()
        r = orig_rotate(res_type, tors)
        row = np.full(5, np.nan)
        row[:len(measured)] = measured[:5]
        rec["ori"].append(row)
        return r

    def choose_rec(weights):
        idx = orig_choose(weights)
        rec["weights"].append(np.array(weights))
        rec["selected"].append(int(idx))
        return idx

    def wrap_efunc(level, f):
        if f is None:
            return None

        def g(new, old):
            e = f(new, old)
            if level > 0:
                rec["steps"].append(dict(level=level, trial=np.array(new, dtype=np.float32),
                                         n_env=old.shape[1], energies=np.array(e)))
            else:
                rec["e_backbone"] = float(e[0])
            return e
        return g

    scb.rotate_sidechain = rotate_rec
    lsb.calc_torsion_angles = torsion_rec
    scb.choose_random = choose_rec
    scb.energy_calculators[:] = [wrap_efunc(i, f) for i, f in enumerate(orig_efuncs)]
    pristine = copy.deepcopy(scb._rotamer_library._data)
    beta = 1 / (300 * 0.008314462)
    out = dict(pdb=np.array(pdb_text), terms=np.array(terms), beta=np.float64(beta),
               n_conf=np.int64(spec["n_conf"]), full=np.bool_(spec["full"]), retain=np.array(retain, dtype=np.int64))
    n_grow = sum(1 for f in orig_efuncs[1:] if f is not None)
    try:
        c = -1
        for attempt in range(spec["n_conf"] if not spec.get("skip_failed") else 12 * spec["n_conf"]):
            if c + 1 >= spec["n_conf"]:
                break
            c += 1
            seed = spec["seed"] * 100 + attempt
            nb_seed(seed)
            uniforms = nb_draw(n_grow)
            nb_seed(seed)
            np.random.seed(seed)
            rec.clear()
            rec.update(chis=[], ori=[], weights=[], selected=[], steps=[])
            structure, ok, energy, _ = scb.create_side_chain_structure([s.coords, beta, retain, None])
            scb._rotamer_library._data = copy.deepcopy(pristine)
            if spec.get("skip_failed") and not ok:          # full-size cases keep conformers that grew to the end
                print("[%s] seed %d died after %d steps: skipped" % (name, seed, len(rec["steps"])))
                c -= 1
                continue
            out["c%d_seed" % c] = np.int64(seed)
            # regroup the flat chi rows per growth step
            pos = 0
            for k, st in enumerate(rec["steps"]):
                T = st["trial"].shape[0]
                chis = np.stack(rec["chis"][pos:pos + T])
                ori = np.stack(rec["ori"][pos:pos + T])
                pos += T
                pre = "c%d_s%d_" % (c, k)
                out[pre + "level"] = np.int64(st["level"])
                out[pre + "chis"] = chis
                out[pre + "ori"] = ori          # measured dihedral (rad, float32 value) before each applied rotation, nan-padded
                out[pre + "energies"] = st["energies"]
                out[pre + "n_env"] = np.int64(st["n_env"])
                if spec["full"]:
                    out[pre + "trial"] = st["trial"]
                if k < len(rec["selected"]):
                    out[pre + "selected"] = np.int64(rec["selected"][k])
                    out[pre + "weights"] = rec["weights"][k]
                    out[pre + "u"] = np.float64(uniforms[k])
            assert pos == len(rec["chis"])
            out["c%d_nsteps" % c] = np.int64(len(rec["steps"]))
            out["c%d_ok" % c] = np.bool_(ok)
            out["c%d_e_backbone" % c] = np.float64(rec.get("e_backbone", np.nan))
            if ok:
                out["c%d_energy" % c] = np.float64(energy)
                # final coordinates: 3-decimal table in residue-number order (what callers get)
                out["c%d_final_coords" % c] = structure.coords
                out["c%d_final_names" % c] = np.array(structure.atom_labels)
                out["c%d_final_resnums" % c] = np.array(structure.res_nums)
            print("[%s] conformer %d ok=%s E=%s steps=%d" % (name, c, ok, energy, len(rec["steps"])))
    finally:
        scb.rotate_sidechain = orig_rotate
        lsb.calc_torsion_angles = orig_torsion
        scb.choose_random = orig_choose
        scb.energy_calculators[:] = orig_efuncs
        scb._rotamer_library._data = pristine
    path = Path(out_dir) / ("%s.npz" % name)
    np.savez_compressed(path, **out)
    print("[%s] wrote %s (%.1f kB)" % (name, path, path.stat().st_size / 1e3))


def make_template_fixture(out_dir):
    """Known answers for every residue template: chi setting and rigid placement."""
    load_reference(0)
    from mcsce.core.build_definitions import sidechain_templates
    from mcsce.libs.libcalc import place_sidechain_template
    from mcsce.libs.libscbuild import rotate_sidechain

    rng = np.random.default_rng(2024)
    out = {}
    for res in sorted(sidechain_templates):
        tmpl, sc_idx = sidechain_templates[res]
        out[res + "_names"] = np.array(tmpl.data_array[:, 2])
        out[res + "_coords"] = tmpl.coords
        out[res + "_sc_idx"] = np.array(sc_idx)
        for n in range(1, 7):
            if res in ("ALA", "GLY"):
                break                                        # no chi: rigid placement only
            tors = rng.uniform(-180, 180, size=(3, n))
            try:
                rot = np.stack([rotate_sidechain(res, t)[0] for t in tors])
            except Exception:
                break
            out["%s_n%d_tors" % (res, n)] = tors
            out["%s_n%d_rot" % (res, n)] = rot
        qr = rng.normal(size=(3, 3))
        rot_m, _ = np.linalg.qr(qr)                      # random orthogonal frame (not aligned with the template)
        if np.linalg.det(rot_m) < 0:
            rot_m[:, 0] = -rot_m[:, 0]
        bb = (tmpl.coords[:3].astype(np.float64) @ rot_m.T + rng.normal(size=3) * 20).astype(np.float32)
        out[res + "_bb"] = bb
        out[res + "_placed"] = place_sidechain_template(bb, tmpl.coords)
    path = Path(out_dir) / "templates.npz"
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f kB)" % (path, path.stat().st_size / 1e3))


if __name__ == "__main__":
    which = sys.argv[1:] or (["templates"] + list(CASES))
    for name in which:
        if name == "templates":
            make_template_fixture(HERE)
        else:
            run_case(name, CASES[name], HERE)
