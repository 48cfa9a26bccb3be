"""One fused growth with the hydrophobic PMF (for ncu captures).  usage: one_hpmf_growth.py [workload] [conformers]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("MCSCE_B200_ALLOW_SYNTHETIC", "1")
import torch, bench
from mcsce_b200.engine import GrowthEngine
from mcsce_b200.hpmf import calculator_for_labels
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
s, sysm, n_conf, mode, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
calc = calculator_for_labels([str(a) for a in sysm.names], [int(r) for r in sysm.resnums], [str(r) for r in sysm.resnames])
out = eng.grow(C, bench.BETA, seed=1, hpmf=calc)
torch.cuda.synchronize()
print("ok", float((out["ok"] == 1).float().mean()))
