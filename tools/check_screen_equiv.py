"""Screen on/off give bit-identical growths on the big workloads (same forced environment split)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, bench, torch
from mcsce_b200.engine import GrowthEngine
for wl, C, split in (("cfg4", 512, 2), ("cfg5", 24, 8), ("cfg3", 1500, 0)):
    s, sysm, n_conf, desc = bench.build_workload(wl)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    a = eng.grow(C, bench.BETA, seed=5, env_split=split, no_graph=True)
    a = {k: a[k].cpu().numpy().copy() for k in ("coords", "energy", "ok")}
    b = eng.grow(C, bench.BETA, seed=5, env_split=split, no_graph=True, no_screen=True)
    same = all(np.array_equal(a[k], b[k].cpu().numpy(), equal_nan=True) for k in ("coords", "energy", "ok"))
    print(wl, C, "identical:", same, "grown %.2f" % a["ok"].mean(), flush=True)
    del eng
