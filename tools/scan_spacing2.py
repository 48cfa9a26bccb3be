import sys
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
eng = GrowthEngine(0)
for cfg in (3, 4, 5, 2):
    for sp in ((16.0, 18.0, 20.0, 24.0) if cfg != 2 else (11.0,)):
        s = Structure(synth.config_backbone(cfg, spacing=sp)).build().remove_side_chains([])
        sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
        eng.set_system(sysm).set_backbone(s.coords)
        n = 256 if cfg != 5 else 16
        out = eng.grow(n, BETA, seed=1, dump=True)
        ok = out["ok"].cpu().numpy(); sel = out["selected"].cpu().numpy()
        grow = [st.index for st in sysm.steps if st.kind == "grow"]
        depth = [(sel[c, grow] >= 0).sum() / len(grow) for c in range(n)]
        print("cfg", cfg, "spacing", sp, "ok %.3f" % ok.mean(), "mean depth %.2f" % np.mean(depth), flush=True)
