"""Tile fragmentation of the scoring pass after the screen: lanes and warps doing work in the CTAs that run."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, bench
from mcsce_b200.engine import GrowthEngine
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
s, sysm, n_conf, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
out = eng.grow(64, bench.BETA, seed=1, dump=True)
te = out["trial_energy"].cpu().numpy()
tmpl_fix = {}
tot_atoms = tot_rows = tot_ctas = 0; w_atoms = w_rows = w_cta = 0.0
hist = np.zeros(9)
from mcsce_b200.chem.tables import load_tables
t = load_tables()
for st in sysm.steps:
    if st.kind != "grow": continue
    o = eng.trial_off[st.index]; T = len(st.prob)
    n_env = int(sysm.cls_active[st.index].sum())
    blk = te[:, o:o+T]
    alive = ~np.isnan(blk).all(axis=1)
    keep = np.isfinite(blk).sum(axis=1)[alive]
    tm = t.templates[st.resname]
    # chi-dependent atoms: moved by some rotation, except the first pivot
    mv = set()
    for prog in tm.chis[:tm.n_sound]:
        mv |= set(int(x) for x in prog["moving"]) if isinstance(prog, dict) else set()
    n_var = getattr(st, "n_var", None)
    if n_var is None:
        n_var = max(1, st.n_sc - 2)      # HA + CB fixed (approximation)
    n_fix = st.n_sc - n_var
    for k in keep:
        atoms = int(k) * n_var + n_fix
        ctas = max(1, -(-atoms // 256))
        for c in range(ctas):
            a = min(256, atoms - c * 256)
            # row-major layout: row r = atoms r*128..; warp w has lanes in row r if r*128 + 32w < a
            rows = sum(1 for r in range(2) for w in range(4) if r * 128 + 32 * w < a)
            tot_atoms += a; tot_rows += rows; tot_ctas += 1
            w_atoms += a * n_env; w_rows += rows * n_env; w_cta += n_env
            hist[rows] += n_env
print(wl, "lane utilisation (atoms / warp-rows*32): %.3f   warp-row utilisation (rows / CTAs*8): %.3f" % (w_atoms / (w_rows * 32), w_rows / (w_cta * 8)))
print("work-weighted share of CTAs by active warp-rows (0..8):", np.round(hist / hist.sum(), 3))
