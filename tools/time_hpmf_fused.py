"""Cost of the hydrophobic PMF inside the fused growth: ms per ensemble with and without the term, per-class kernel times."""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("MCSCE_B200_ALLOW_SYNTHETIC", "1")
import torch, bench
from mcsce_b200.engine import GrowthEngine
from mcsce_b200.hpmf import calculator_for_labels
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [256, 2048]
s, sysm, n_conf, mode, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
calc = calculator_for_labels([str(a) for a in sysm.names], [int(r) for r in sysm.resnums], [str(r) for r in sysm.resnames])
rows = []
for C in counts:
    res = {}
    for tag, kw in (("plain", {}), ("hpmf", dict(hpmf=calc))):
        for k in range(2): eng.grow(C, bench.BETA, seed=k, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for k in range(3): out = eng.grow(C, bench.BETA, seed=10 + k, **kw)
        e1.record(); torch.cuda.synchronize()
        eng.profile_begin(); eng.grow(C, bench.BETA, seed=99, **kw); prof = eng.profile_end()
        res[tag] = dict(ms=e0.elapsed_time(e1) / 3, ok=float((out["ok"] == 1).float().mean()), kernel_ms={k: round(v[0], 2) for k, v in prof.items()})
    res.update(workload=wl, conformers=C, ratio=res["hpmf"]["ms"] / res["plain"]["ms"])
    rows.append(res)
    print("%s C=%d: plain %.1f ms, with hpmf %.1f ms (x%.2f); kernels with the term %s" % (wl, C, res["plain"]["ms"], res["hpmf"]["ms"], res["ratio"], res["hpmf"]["kernel_ms"]), flush=True)
print(json.dumps(rows))
