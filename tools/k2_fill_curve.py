"""Cost of a K2 scoring launch vs how full its tiles are: one mid-protein residue of cfg3, 2048 environments, T trials each
(no dedup, no screen: T x n_sc atoms per environment).  Shows whether a tile's time follows its atoms or is fixed."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, bench, torch
from mcsce_b200.engine import GrowthEngine
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
s, sysm, n_conf, desc = bench.build_workload("cfg3")
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
out = eng.grow(8, bench.BETA, seed=2)
c = int(torch.argmax(out["ok"]).item())
coords = out["coords"][c]
res = [st for st in sysm.steps if st.kind == "grow" and st.resname == (sys.argv[2] if len(sys.argv) > 2 else "PHE") and st.index > 120][0]
print("residue", res.index, res.resname, "n_sc", res.n_sc, "environment atoms", res.sc_start)
old = coords[:res.sc_start].unsqueeze(0).expand(S, -1, -1).contiguous()
chis = np.repeat(res.chi_mean[:1], 64, axis=0) + np.random.default_rng(0).normal(0, 3.0, (64, res.chi_mean.shape[1]))
tx = eng.place_trials(res.index, chis)                          # (64, n_sc, 3)
Ts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8, 12, 16, 17, 18, 24, 34, 51, 64]
for T in Ts:
    trial = tx[:T].unsqueeze(0).expand(S, -1, -1, -1).contiguous()
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        eng.score_trials(res.index, trial, old, env_split=1)
        e1.record(); torch.cuda.synchronize()
    atoms = T * res.n_sc
    print("T %3d atoms/env %4d tiles %d: %.3f ms  (%.2f ns per atom-environment)" % (T, atoms, -(-atoms // 256), e0.elapsed_time(e1), e0.elapsed_time(e1) * 1e6 / (S * atoms)))
