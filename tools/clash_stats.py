import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, bench
from mcsce_b200.engine import GrowthEngine
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
s, sysm, n_conf, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
out = eng.grow(32, bench.BETA, seed=1, dump=True)
te = out["trial_energy"].cpu().numpy()
ev = ~np.isnan(te)
print(wl, "evaluated trials", ev.sum(), "clashing fraction %.3f" % (np.isinf(te[ev]).mean()))
# weight by work: pairs per trial ~ n_sc * n_env
w_tot = w_clash = 0.0
for st in sysm.steps:
    if st.kind != "grow": continue
    o = eng.trial_off[st.index]; T = len(st.prob)
    n_env = int(sysm.cls_active[st.index].sum())
    blk = te[:, o:o+T]
    e = ~np.isnan(blk)
    w_tot += e.sum() * st.n_sc * n_env; w_clash += (np.isinf(blk) & e).sum() * st.n_sc * n_env
print("work-weighted clashing fraction %.3f" % (w_clash / w_tot))
