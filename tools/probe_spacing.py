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
for cfg in (3, 4, 5):
    for sp in (11.0, 12.0, 13.0, 14.0, 16.0):
        s = Structure(synth.config_backbone(cfg, spacing=sp)).build().remove_side_chains([])
        sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
        eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
        n = 256 if cfg != 5 else 32
        out = eng.grow(n, BETA, seed=1, dump=True)
        sel = out["selected"].cpu().numpy(); ok = out["ok"].cpu().numpy()
        grow = [st.index for st in sysm.steps if st.kind == "grow"]
        fail_at = []
        for c in range(n):
            if not ok[c]:
                f = [i for i in grow if sel[c, i] < 0]
                fail_at.append(f[0])
        names = [sysm.steps[i].resname for i in fail_at]
        from collections import Counter
        print("cfg", cfg, "spacing", sp, "ok %.3f" % ok.mean(), "E_in", eng.input_energy(), "first-fail residues", Counter(fail_at).most_common(4), Counter(names).most_common(4), flush=True)
        eng.close()
