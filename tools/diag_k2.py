"""Diagnostic: K2 (and K1+K2) error distribution against the golden fixtures. Writes gpurun_out/diag_k2.npz."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from test_gpu_parity import engine_for, oracle_env_sequence, FULL_CASES

out = {}
for name in FULL_CASES:
    case, s, sysm, eng = engine_for(name)
    rows = []
    for c in range(case.n_conf):
        envs = oracle_env_sequence(case, c, s, sysm)
        for i, g in case.steps_by_level(c).items():
            got = eng.score_trials(i, g["trial"][None], envs[i][None]).cpu().numpy()[0]
            ref = g["energies"]
            fin = ~np.isinf(ref)
            for t in np.where(fin)[0]:
                rows.append((c, i, t, ref[t], got[t]))
    a = np.array(rows)
    out[name] = a
    d = np.abs(a[:, 4] - a[:, 3]); tol = np.maximum(1e-3, 1e-5 * np.abs(a[:, 3]))
    r = d / tol
    k = np.argsort(-r)[:5]
    print(name, "n=%d  ratio-to-tol: max %.3f  p99 %.3f  median %.4f ; abs max %.2e" % (len(a), r.max(), np.quantile(r, .99), np.median(r), d.max()))
    for j in k:
        print("    res %d (%s) trial %d  E=%.4f  d=%.2e  ratio %.2f" % (a[j, 1], sysm.steps[int(a[j, 1])].resname, a[j, 2], a[j, 3], d[j], r[j]))
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
np.savez(ROOT / "gpurun_out" / "diag_k2.npz", **out)
