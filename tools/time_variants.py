"""Time K2 variants: for each .so in build_variants/, run cfg3 with C conformers and report per-kernel ms."""
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.grow(C, bench.BETA, seed=1); e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    eng.profile_begin(); eng.grow(C, bench.BETA, seed=1); prof = eng.profile_end()
    print(json.dumps({"total_ms": total, "k2_ms": prof["score_generic"][0], "sp_ms": prof["score_special"][0],
                      "place_ms": prof["place_trials"][0], "ok": float(out["ok"].float().mean())}))
else:
    C = sys.argv[1] if len(sys.argv) > 1 else "1024"
    wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
    for lib in sorted((ROOT / "build_variants").glob("*.so")):
        env = dict(os.environ, MCSCE_B200_LIB=str(lib))
        r = subprocess.run([sys.executable, __file__, "--child", C, wl], env=env, capture_output=True, text=True)
        print(lib.name, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
