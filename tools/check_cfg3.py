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
from oracle import ref_numpy as O
lib, ptm = DunbrakRotamerLibrary(), ptmRotamerLib()
BETA = 1.0 / (300 * 0.008314462)
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
s = Structure(synth.config_backbone(cfg)).build().remove_side_chains([])
sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
C_ = 4
rng = np.random.default_rng(0)
u = rng.random((C_, sysm.n_res))
out = eng.grow(C_, BETA, seed=5, uniforms=u, dump=True)
chis = out["chis"].cpu().numpy(); sel = out["selected"].cpu().numpy(); te = out["trial_energy"].cpu().numpy()
print("ok", out["ok"].cpu().numpy(), "first -1:", [next((i for i in range(sysm.n_res) if sel[c, i] == -1 and sysm.steps[i].kind == "grow"), None) for c in range(C_)])
for c in range(2):
    def chi_for_step(i, res, phi, psi, c=c):
        st = sysm.steps[i]; o = eng.trial_off[i]
        return chis[c, o:o + len(st.prob), :st.n_chi_lib], st.prob
    rec = []
    ref = O.grow_conformer(s, s.coords, BETA, chi_for_step, lambda i, c=c: u[c, i], record=rec)
    print("oracle conf", c, "ok", ref["ok"], "failed_at", ref.get("failed_at"))
    for r in rec:
        i = r["step"]; o = eng.trial_off[i]
        got = te[c, o:o + len(r["energies"])]
        same_clash = np.array_equal(np.isinf(got), np.isinf(r["energies"]))
        fin = ~np.isinf(r["energies"]) & ~np.isinf(got)
        err = np.abs(got[fin] - r["energies"][fin]).max() if fin.any() else 0
        flag = "" if same_clash and ("selected" not in r or sel[c, i] == r["selected"]) else "   <<<<<< MISMATCH"
        print("  step %d %s T=%d clash gpu=%d ref=%d maxerr=%.2e sel gpu=%d ref=%s%s" % (i, r["res"], len(got), np.isinf(got).sum(), np.isinf(r["energies"]).sum(), err, sel[c, i], r.get("selected"), flag))
        if flag: break
