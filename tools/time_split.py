"""Growth time vs forced environment split (cfg3 by default)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench, torch
from mcsce_b200.engine import GrowthEngine
C = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
s, sysm, n_conf, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
eng.grow(C, bench.BETA, seed=0, env_split=8); torch.cuda.synchronize()
for split in (0, 1, 2, 3, 4, 6, 8):
    for no_screen in (False, True):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.grow(C, bench.BETA, seed=1, env_split=split, no_screen=no_screen)
        e0.record(); out = eng.grow(C, bench.BETA, seed=1, env_split=split, no_screen=no_screen); e1.record(); torch.cuda.synchronize()
        print("split %d screen %d: %.1f ms" % (split, not no_screen, e0.elapsed_time(e1)), flush=True)
