"""Growth time vs number of conformer groups in flight (MCSCE_B200_LANES)."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, str(ROOT))
    import bench, torch
    from mcsce_b200.engine import GrowthEngine
    C = int(sys.argv[2]); wl = sys.argv[3]
    s, sysm, n_conf, desc = bench.build_workload(wl)
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    eng.grow(C, bench.BETA, seed=0); torch.cuda.synchronize()
    ts = []
    for k in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = eng.grow(C, bench.BETA, seed=1 + k); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%.1f ms (min of 3: %s), esum %.6f" % (min(ts), [round(t, 1) for t in ts], float(out["energy"][out["ok"].bool()].sum())))
else:
    C = sys.argv[1] if len(sys.argv) > 1 else "2048"
    wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
    for g in ("1", "2", "3", "4"):
        r = subprocess.run([sys.executable, __file__, "--child", C, wl], env=dict(os.environ, MCSCE_B200_LANES=g), capture_output=True, text=True)
        print("lanes", g, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
