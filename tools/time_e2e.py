import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench, torch
from mcsce_b200.engine import GrowthEngine
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
s, sysm, n_conf, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng.grow(C, bench.BETA, seed=0); torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); eng.set_backbone(s.coords); torch.cuda.synchronize(); t1 = time.perf_counter()
    out = eng.grow(C, bench.BETA, seed=rep); torch.cuda.synchronize(); t2 = time.perf_counter()
    out = eng.grow(C, bench.BETA, seed=rep, host_out=True); t3 = time.perf_counter()
    print("set_backbone %.1f ms | grow(device) %.1f ms | grow(host_out) %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
eng.profile_begin(); eng.grow(C, bench.BETA, seed=9); print(eng.profile_end())
