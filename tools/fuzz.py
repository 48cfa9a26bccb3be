"""Randomised consistency run: random sequences (standard + PTM residues), sizes and conformer counts.
For every case: (1) screen on/off and dedup on/off give identical growths (same forced split), (2) the C oracle
replaying the GPU's chi rows and uniforms makes the same selections and clash decisions."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from golden_utils import library
from mcsce_b200 import synth
from mcsce_b200.engine import GrowthEngine
from mcsce_b200.structure import Structure
from mcsce_b200.system import build_system
from oracle.growth_c import COracle
BETA = 1.0 / (300 * 0.008314462)
lib, ptm = library()
STD = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIP", "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"]
PTM = ["SEP", "TPO", "PTR", "H2D", "H2E", "M3L", "KCX"]


def make_case(seed):
    """Random bundle, engine and growth arguments of fuzz case `seed`."""
    rng = np.random.default_rng(seed)
    n_hel = int(rng.integers(1, 4)); L = int(rng.integers(8, 45))
    seqs = []
    for h in range(n_hel):
        seq = [str(rng.choice(STD)) for _ in range(L)]
        for p in range(L):
            if 0 < p < L - 1 and rng.random() < 0.08: seq[p] = str(rng.choice(PTM))     # the force field has no terminal PTM residues
            if seq[p] == "PRO" and p > 2: seq[p] = "ALA"
        if seq[0] in ("PRO",): seq[0] = "ALA"
        seqs.append(seq)
    pdb = synth.helix_bundle_pdb(seqs, seed=int(rng.integers(1 << 30)), one_chain=bool(rng.integers(2)) or n_hel == 1, spacing=float(rng.uniform(11, 16)))
    s = Structure(pdb).build().remove_side_chains([])
    sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    C = int(rng.choice([1, 3, 17, 64]))
    split = int(rng.choice([1, 2, 5]))
    u = rng.random((C, sysm.n_res))
    sd = int(rng.integers(1 << 30))
    return dict(s=s, sysm=sysm, eng=eng, C=C, split=split, u=u, sd=sd, n_hel=n_hel, L=L)


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    t0 = time.time()
    for case in range(n_cases):
        k = make_case(seed0 + case)
        s, sysm, eng, C, split, u, sd, n_hel, L = (k[x] for x in ("s", "sysm", "eng", "C", "split", "u", "sd", "n_hel", "L"))
        a = eng.grow(C, BETA, seed=sd, uniforms=u, dump=True, no_graph=True, env_split=split)
        a = {k2: a[k2].cpu().numpy().copy() for k2 in ("coords", "energy", "ok", "selected", "trial_energy", "chis")}
        msgs = []
        for kw in (dict(no_screen=True), dict(no_dedup=True, no_screen=True), dict(no_graph=2), dict(no_graph=3)):
            args = dict(no_graph=True); args.update(kw)
            b = eng.grow(C, BETA, seed=sd, uniforms=u, dump=True, env_split=split, **args)
            exact = "no_dedup" not in kw and kw.get("no_graph") != 3      # (3: persistent kernel, other float32 grouping)
            if not np.array_equal(a["selected"], b["selected"].cpu().numpy()) or not np.array_equal(a["ok"], b["ok"].cpu().numpy()):
                msgs.append("selection/ok differs with %s" % kw)
            elif exact and not np.array_equal(a["energy"], b["energy"].cpu().numpy(), equal_nan=True):
                msgs.append("energy not bit-identical with %s" % kw)
        co = COracle(sysm, s.coords)
        ref = co.grow(C, BETA, chis=a["chis"], uniforms=u, dump=True)
        soft = []            # differences from the oracle that K1's float32 noise can cause (borderline clash / draw)
        if not np.array_equal(ref["ok"], a["ok"]): soft.append("ok differs from oracle")
        for c in range(C):
            for st in sysm.steps:
                if st.kind != "grow": continue
                if a["selected"][c, st.index] != ref["selected"][c, st.index]:
                    soft.append("selection differs from oracle: conf %d residue %d %s" % (c, st.index, st.resname)); break
                if a["selected"][c, st.index] < 0: break
        te, tr = a["trial_energy"], ref["trial_energy"]
        both = ~np.isnan(te) & ~np.isnan(tr)
        if not np.array_equal(np.isinf(te[both]), np.isinf(tr[both])): soft.append("clash flags differ from oracle")
        print("case %d: %d helices x %d residues, %d atoms, C=%d split=%d grown %.2f %s%s" % (
            seed0 + case, n_hel, L, sysm.n_atoms, C, split, a["ok"].mean(), "OK" if not msgs else "FAIL: " + "; ".join(msgs),
            "" if not soft else "  [vs oracle: " + "; ".join(soft[:2]) + "]"), flush=True)
        bad += bool(msgs)
        del eng
    print("%d cases, %d failed, %.0f s" % (n_cases, bad, time.time() - t0))


if __name__ == "__main__":
    main()
