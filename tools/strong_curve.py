"""One-GPU view of the strong-scaling curve: milliseconds per ensemble of cfg3 (or another workload) for the conformer
counts one GPU holds when 2048 conformers are sharded over 1..32 GPUs.  usage: strong_curve.py [workload] [counts]"""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("MCSCE_B200_ALLOW_SYNTHETIC", "1")
import torch, bench
from mcsce_b200.engine import GrowthEngine
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2048, 1024, 512, 256, 128, 64]
SPLIT = int(os.environ.get("SPLIT", "0"))
s, sysm, n_conf, mode, desc = bench.build_workload(wl)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
rows = []
for C in counts:
    for k in range(2): eng.grow(C, bench.BETA, seed=k, env_split=SPLIT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(3): eng.grow(C, bench.BETA, seed=10 + k, env_split=SPLIT)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    eng.counters(reset=True); eng.profile_begin(); eng.grow(C, bench.BETA, seed=99, env_split=SPLIT); prof = eng.profile_end()
    _, pa, pc = eng.counters(reset=True, computed=True)
    k2 = prof["score_generic"][0]
    frac = 22 * pc / (k2 * 1e-3) / 1e12 / 74.45 if k2 else 0
    rows.append(dict(workload=wl, conformers=C, ms=ms, conf_per_s=C / ms * 1e3, us_per_conformer=ms / C * 1e3, k2_ms=k2, frac_executed=frac,
                     kernel_ms={k: v[0] for k, v in prof.items()}))
    print("%s C=%5d: %8.2f ms  %8.1f conf/s  %7.1f us/conf  K2 %.1f ms frac_executed %.3f  others %s" % (
        wl, C, ms, C / ms * 1e3, ms / C * 1e3, k2, frac, {k: round(v[0], 1) for k, v in prof.items() if k != "score_generic"}), flush=True)
print(json.dumps(rows))
