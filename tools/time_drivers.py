"""Milliseconds per ensemble for the three growth drivers (mcsce_grow_opts.no_graph: 1 per-kernel path, 2 replayed CUDA graph,
3 persistent cluster kernel) on a workload at several conformer counts.
usage: time_drivers.py [workload] [counts] [modes]      e.g.  time_drivers.py cfg2 1,16,128 1,2,3"""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("MCSCE_B200_ALLOW_SYNTHETIC", "1")
import numpy as np, torch, bench
from mcsce_b200.engine import GrowthEngine
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8, 32, 128]
modes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 3]
REPS = int(os.environ.get("REPS", "5"))
s, sysm, n_conf, mode, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
rows = []
for C in counts:
    for m in modes:
        kw = dict(no_graph=m)
        if m == 1 and os.environ.get("NO_SCREEN"): kw["no_screen"] = True
        for k in range(2): out = eng.grow(C, bench.BETA, seed=k, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for k in range(REPS): out = eng.grow(C, bench.BETA, seed=10 + k, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / REPS
        eng.counters(reset=True); out = eng.grow(C, bench.BETA, seed=99, **kw); torch.cuda.synchronize()
        launches, pa, pc = eng.counters(reset=True, computed=True)
        okf = float(out["ok"].float().mean())
        rows.append(dict(workload=wl, conformers=C, driver=m, ms=ms, conf_per_s=C / ms * 1e3, launches=launches, pairs=pa, pairs_executed=pc,
                         pairs_per_s=pa / ms * 1e3, executed_frac_of_fp32_peak=22 * pc / (ms * 1e-3) / 1e12 / 74.45, grown=okf))
        print("%s C=%5d driver %d: %9.3f ms  %9.1f conf/s  launches %5d  executed pairs/s %.3e (%.3f of FP32 peak incl. everything)  grown %.2f" % (
            wl, C, m, ms, C / ms * 1e3, launches, pc / ms * 1e3, rows[-1]["executed_frac_of_fp32_peak"], okf), flush=True)
print(json.dumps(rows))
