import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, bench, collections
from mcsce_b200.engine import GrowthEngine
s, sysm, n_conf, desc = bench.build_workload("cfg3")
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
out = eng.grow(128, bench.BETA, seed=1, dump=True)
te = out["trial_energy"].cpu().numpy()
by = collections.defaultdict(lambda: [0.0, 0.0])
for st in sysm.steps:
    if st.kind != "grow": continue
    o = eng.trial_off[st.index]; T = len(st.prob)
    blk = te[:, o:o+T]; alive = ~np.isnan(blk).all(axis=1)
    by[st.resname][0] += np.isfinite(blk[alive]).sum(); by[st.resname][1] += alive.sum() * T
print({k: round(a / b, 3) for k, (a, b) in sorted(by.items())})
