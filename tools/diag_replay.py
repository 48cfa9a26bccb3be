"""Which trials differ most between the GPU growth and the C oracle replaying the same chi rows / uniforms."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, bench
from mcsce_b200.engine import GrowthEngine
from oracle.growth_c import COracle
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
s, sysm, n_conf, mode, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
co = COracle(sysm, s.coords)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 11
u = np.random.default_rng(5).random((n, sysm.n_res))
out = eng.grow(n, bench.BETA, seed=seed, uniforms=u, dump=True)
ref = co.grow(n, bench.BETA, chis=out["chis"].cpu().numpy(), uniforms=u, dump=True)
te, tr = out["trial_energy"].cpu().numpy(), ref["trial_energy"]
fin = ~np.isnan(te) & ~np.isnan(tr) & np.isfinite(tr) & np.isfinite(te)
ratio = np.where(fin, np.abs(te - tr) / np.maximum(1e-3, 1e-5 * np.abs(tr)), 0)
offs = np.cumsum([0] + [len(x.prob) if x.kind == "grow" else 0 for x in sysm.steps])
order = np.argsort(ratio.ravel())[::-1][:25]
for f in order:
    c, t = divmod(int(f), te.shape[1])
    i = int(np.searchsorted(offs, t, side="right") - 1)
    print("ratio %.2f conf %d step %d %s trial %d  gpu %.6f  oracle %.6f  diff %.2e" % (ratio[c, t], c, i, sysm.steps[i].resname, t - offs[i], te[c, t], tr[c, t], te[c, t] - tr[c, t]))
by = {}
for i, st in enumerate(sysm.steps):
    if st.kind != "grow": continue
    r = ratio[:, offs[i]:offs[i+1]]
    by.setdefault(st.resname, []).append(r.max())
print({k: round(float(np.max(v)), 2) for k, v in sorted(by.items())})
