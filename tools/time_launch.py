import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench, torch
from mcsce_b200.engine import GrowthEngine
wl = sys.argv[1]; C = int(sys.argv[2])
s, sysm, n_conf, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
for fuse in (False, True):   # here: False = direct launches, True = CUDA graph replay
    eng.grow(C, bench.BETA, seed=0, no_graph=not fuse); torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter(); out = eng.grow(C, bench.BETA, seed=rep, no_graph=not fuse); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print("%s C=%d graph=%s: host enqueue %.2f ms, total %.2f ms" % (wl, C, fuse, (t1 - t0) * 1e3, (t2 - t0) * 1e3), flush=True)
