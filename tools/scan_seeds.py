import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from mcsce_b200 import synth
from mcsce_b200.core.rotamer_library import DunbrakRotamerLibrary, ptmRotamerLib
from mcsce_b200.structure import Structure
from mcsce_b200.system import build_system
from mcsce_b200.engine import GrowthEngine
lib, ptm = DunbrakRotamerLibrary(), ptmRotamerLib()
BETA = 1.0 / (300 * 0.008314462)
L = int(sys.argv[1]); nh = int(sys.argv[2]); seeds = range(int(sys.argv[3]), int(sys.argv[4])); sp = float(sys.argv[5])
eng = GrowthEngine(0)
for seed in seeds:
    seq = synth.random_sequence(L, seed)
    hl = L // nh
    pdb = synth.helix_bundle_pdb([seq[i:i + hl] for i in range(0, L, hl)], seed=seed, one_chain=True, spacing=sp)
    s = Structure(pdb).build().remove_side_chains([])
    sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
    eng.set_system(sysm).set_backbone(s.coords)
    out = eng.grow(256, BETA, seed=1, dump=True)
    ok = out["ok"].cpu().numpy(); sel = out["selected"].cpu().numpy()
    grow = [st.index for st in sysm.steps if st.kind == "grow"]
    depth = [(sel[c, grow] >= 0).sum() / len(grow) for c in range(256)]
    print("L", L, "helices", nh, "seed", seed, "spacing", sp, "ok %.3f" % ok.mean(), "mean depth %.2f" % np.mean(depth), flush=True)
