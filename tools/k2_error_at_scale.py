"""K2 + K2s error on identical coordinates at full size: the energy the growth scored for the trial it selected vs the
C oracle's float64 re-scoring of the same residue on the GPU's own final coordinates (no K1 involved)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, bench
from mcsce_b200.engine import GrowthEngine
from oracle.growth_c import COracle
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
s, sysm, n_conf, mode, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
co = COracle(sysm, s.coords)
out = eng.grow(n, bench.BETA, seed=7, dump=True, no_graph=True)
te, sel, ok, xyz = (out[k].cpu().numpy() for k in ("trial_energy", "selected", "ok", "coords"))
rs = []; by = {}
for c in range(n):
    if not ok[c]: continue
    _, ratio, lev = co.rescore_levels(xyz[c])
    for st in sysm.steps:
        if st.kind != "grow": continue
        e_gpu = te[c, eng.trial_off[st.index] + sel[c, st.index]]
        e_ref = lev[st.index]
        r = abs(e_gpu - e_ref) / max(1e-3, 1e-5 * abs(e_ref))
        rs.append(r); by.setdefault(st.resname, []).append(r)
rs = np.array(rs)
print(wl, "%d residues of %d conformers: |dE| / tolerance  max %.3f  p99 %.3f  median %.4f" % (len(rs), int(ok.sum()), rs.max(), np.quantile(rs, .99), np.median(rs)))
print({k: round(float(np.max(v)), 2) for k, v in sorted(by.items())})
