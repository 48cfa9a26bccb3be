"""Sweep the neighbour-screen radius scale (MCSCE_B200_SCREEN_SCALE; 0 = off): growth time, K2 time, executed pairs."""
import os, subprocess, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, str(ROOT))
    import bench, torch
    from mcsce_b200.engine import GrowthEngine
    C = int(sys.argv[2]); wl = sys.argv[3]
    s, sysm, n_conf, desc = bench.build_workload(wl)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    eng.grow(C, bench.BETA, seed=0)
    torch.cuda.synchronize()
    eng.counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.grow(C, bench.BETA, seed=1); e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    _, alg, done = eng.counters(reset=True, computed=True)
    eng.profile_begin(); eng.grow(C, bench.BETA, seed=1); prof = eng.profile_end()
    print(json.dumps({"total_ms": round(total, 1), "k2_ms": round(prof["score_generic"][0], 1), "sp_ms": round(prof["score_special"][0], 1),
                      "place_ms": round(prof["place_trials"][0], 1), "sel_ms": round(prof["select"][0], 1),
                      "executed/algorithmic": round(done / alg, 3), "ok": float(out["ok"].float().mean()),
                      "esum": float(out["energy"][out["ok"].bool()].sum())}))
else:
    C = sys.argv[1] if len(sys.argv) > 1 else "1024"
    wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
    scales = sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "0.5", "0.65", "0.8", "1.0"]
    for sc in scales:
        env = dict(os.environ, MCSCE_B200_SCREEN_SCALE=sc)
        r = subprocess.run([sys.executable, __file__, "--child", C, wl], env=env, capture_output=True, text=True)
        print("scale", sc, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:], flush=True)
