"""Tile occupancy of the scoring pass after the screen: how full the CTAs that run are (cfg3 by default)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, bench
from mcsce_b200.engine import GrowthEngine
from mcsce_b200.chem.tables import load_tables
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
s, sysm, n_conf, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
out = eng.grow(64, bench.BETA, seed=1, dump=True)
te = out["trial_energy"].cpu().numpy()
t = load_tables()
hist_after = np.zeros(5); hist_before = np.zeros(5)       # work-weighted CTAs by warps in use (64 atoms per warp)
cta_after = cta_before = 0.0; pairs_after = pairs_before = 0.0
for st in sysm.steps:
    if st.kind != "grow": continue
    o = eng.trial_off[st.index]; T = len(st.prob)
    n_env = int(sysm.cls_active[st.index].sum())
    blk = te[:, o:o+T]
    alive = ~np.isnan(blk).all(axis=1)
    keep = np.isfinite(blk).sum(axis=1)[alive]
    tm = t.templates[st.resname]
    moved = set()
    for prog in tm.chis[:min(tm.n_sound, st.n_chi)]:
        moved |= set(int(x) for x in prog["moving"])
    if st.n_chi > 0: moved.discard(int(tm.chis[0]["axis"][1]))
    n_var = sum(1 for p in tm.sc_idx if int(p) in moved); n_fix = st.n_sc - n_var
    for label, ks in (("after", keep), ("before", np.full(len(keep), T))):
        for k in ks:
            atoms = int(k) * n_var + n_fix
            full, rem = divmod(atoms, 256)
            tiles = [256] * full + ([rem] if rem else [])
            for a in tiles:
                w = -(-a // 64)
                if label == "after": hist_after[w] += n_env; cta_after += n_env; pairs_after += a * n_env
                else: hist_before[w] += n_env; cta_before += n_env; pairs_before += a * n_env
print(wl, "CTA-time share by warps in use (1..4): before", np.round(hist_before[1:] / hist_before.sum(), 3), " after", np.round(hist_after[1:] / hist_after.sum(), 3))
print("mean atoms per running CTA: before %.0f after %.0f;  CTA-time after/before %.3f; pairs after/before %.3f" % (
    pairs_before / cta_before, pairs_after / cta_after, cta_after / cta_before, pairs_after / pairs_before))
