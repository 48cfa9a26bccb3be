"""Look into one fuzz case whose clash flags differ from the C oracle: how close to its clash distance is the pair?"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
sys.path.insert(0, str(ROOT / "tools"))
import fuzz
case = int(sys.argv[1])
k = fuzz.make_case(case)
sysm, eng, s, C, split, u, sd = (k[x] for x in ("sysm", "eng", "s", "C", "split", "u", "sd"))
BETA = fuzz.BETA
from oracle.growth_c import COracle
a = eng.grow(C, BETA, seed=sd, uniforms=u, dump=True, no_graph=True, env_split=split)
te = a["trial_energy"].cpu().numpy(); chis = a["chis"].cpu().numpy(); sel = a["selected"].cpu().numpy()
co = COracle(sysm, s.coords)
ref = co.grow(C, BETA, chis=chis, uniforms=u, dump=True)
tr = ref["trial_energy"]
both = ~np.isnan(te) & ~np.isnan(tr)
mism = np.argwhere(both & (np.isinf(te) != np.isinf(tr)))
print("case", case, "mismatching clash flags:", len(mism))
offs = eng.trial_off; nts = eng.n_trials
rad_thr = sysm.cls_thr
for c, t in mism[:5]:
    i = int(np.searchsorted(offs + (nts == 0) * 10**9, t, side="right") - 1)
    i = max(j for j in range(sysm.n_res) if nts[j] > 0 and offs[j] <= t)
    st = sysm.steps[i]; tl = t - offs[i]
    one = co.grow(1, BETA, chis=chis[c:c+1], uniforms=u[c:c+1], dump=True)
    x_or = one["trial_xyz"][t, :st.n_sc]
    x_gpu = eng.place_trials(i, chis[c, t][None])[0].cpu().numpy()[:, :3] if hasattr(eng, "place_trials") else None
    env = one["coords"][0]; placed = ~np.isnan(env[:, 0]); placed[st.sc_start:] = False
    sp = set(int(q) for q in st.special_env)
    best = {}
    for name, x in (("oracle", x_or), ("gpu", x_gpu)):
        if x is None: continue
        m = 1e9
        for k in range(st.n_sc):
            ck = sysm.cls_of_atom[st.sc_start + k]
            for j in np.where(placed)[0]:
                if int(j) in sp: continue
                thr = rad_thr[ck, sysm.cls_of_atom[j]]
                if thr > 0:
                    d = np.sqrt(float(np.float32(((x[k] - env[j]).astype(np.float32) ** 2).sum())))
                    m = min(m, d / thr)
        for p_ in range(len(st.sp_k)):                      # special pairs: exact per-pair clash distances
            thr = st.sp_coef[p_, 3]
            if thr <= 0: continue
            kk, pa = int(st.sp_k[p_]), int(st.sp_partner[p_])
            other = env[int(st.special_env[pa])] if pa >= 0 else x[-pa - 1]
            d = np.sqrt(float(np.float32(((x[kk] - other).astype(np.float32) ** 2).sum())))
            m = min(m, d / thr)
        best[name] = m
    print("  conf %d residue %d %s trial %d: gpu E %s oracle E %s | min distance/clash distance: %s | max |dx| %.2e" % (
        c, i, st.resname, tl, te[c, t], tr[c, t], {k: "%.8f" % v for k, v in best.items()}, np.abs(x_or - x_gpu).max() if x_gpu is not None else -1))
