"""Short single-GPU run for ncu: cfg3 backbone, C conformers, one warm-up growth + one growth."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import os
os.environ.setdefault("MCSCE_B200_ALLOW_SYNTHETIC", "1")
import bench
from mcsce_b200.engine import GrowthEngine
import torch
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 512
s, sysm, n_conf, mode, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
for k in range(2):
    out = eng.grow(C, bench.BETA, seed=k)
torch.cuda.synchronize()
print("ok", float(out["ok"].float().mean()), eng.counters())
