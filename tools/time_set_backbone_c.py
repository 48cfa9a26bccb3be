"""C-side share of set_backbone: wall time of the mcsce_set_backbone call alone (cfg2, cfg3, cfg5)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench, torch
from mcsce_b200.engine import GrowthEngine
import mcsce_b200.engine as E
for wl in (sys.argv[1:] or ["cfg2", "cfg3", "cfg5"]):
    s, sysm, n_conf, desc = bench.build_workload(wl)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    lib = eng.lib
    orig = lib.mcsce_set_backbone
    acc = [0.0]
    class W:
        def __call__(self, *a):
            t0 = time.perf_counter(); r = orig(*a); acc[0] += time.perf_counter() - t0; return r
    eng.lib = type("L", (), {"__getattr__": lambda self, k: W() if k == "mcsce_set_backbone" else getattr(lib, k)})()
    t0 = time.perf_counter()
    for _ in range(5): eng.set_backbone(s.coords)
    tot = (time.perf_counter() - t0) / 5
    print(wl, "set_backbone %.2f ms, of which the C call %.2f ms" % (tot * 1e3, acc[0] / 5 * 1e3), flush=True)
    acc[0] = 0.0
    t0 = time.perf_counter(); g = 0.0
    for k in range(5):
        eng.set_backbone(s.coords)
        tg = time.perf_counter(); eng.grow(min(n_conf, 512), bench.BETA, seed=k, host_out=True); g += time.perf_counter() - tg
    tot = (time.perf_counter() - t0 - g) / 5
    print(wl, "alternating with growths: set_backbone %.2f ms, of which the C call %.2f ms" % (tot * 1e3, acc[0] / 5 * 1e3), flush=True)
