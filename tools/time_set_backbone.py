"""Host cost of GrowthEngine.set_backbone by part (cProfile, cfg3 and cfg2)."""
import sys, time, cProfile, pstats
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench, torch
from mcsce_b200.engine import GrowthEngine
for wl in (sys.argv[1:] or ["cfg2", "cfg3"]):
    s, sysm, n_conf, desc = bench.build_workload(wl)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    t0 = time.perf_counter()
    for _ in range(10): eng.set_backbone(s.coords)
    torch.cuda.synchronize()
    print(wl, "set_backbone %.2f ms" % ((time.perf_counter() - t0) * 100))
    pr = cProfile.Profile(); pr.enable()
    for _ in range(10): eng.set_backbone(s.coords)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
