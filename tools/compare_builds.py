"""Per-trial energy differences between kernel builds (.so files in build_variants/) on the same seeded cfg3 growth:
the first library is the reference; conformers are compared up to the first residue where the selections differ."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, str(ROOT))
    import numpy as np, bench
    from mcsce_b200.engine import GrowthEngine
    s, sysm, n_conf, desc = bench.build_workload("cfg3")
    eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
    u = np.random.default_rng(1).random((8, sysm.n_res))
    out = eng.grow(8, bench.BETA, seed=3, uniforms=u, dump=True, no_graph=True, env_split=1)
    np.savez(sys.argv[2], te=out["trial_energy"].cpu().numpy(), sel=out["selected"].cpu().numpy(), off=eng.trial_off, nt=eng.n_trials)
else:
    import numpy as np
    libs = sorted((ROOT / "build_variants").glob("*.so"))
    res = []
    for lib in libs:
        f = "/tmp/cmp_%s.npz" % lib.stem
        subprocess.run([sys.executable, __file__, "--child", f], env=dict(os.environ, MCSCE_B200_LIB=str(lib)), check=True, capture_output=True)
        res.append(np.load(f))
    a = res[0]
    for lib, b in zip(libs[1:], res[1:]):
        worst = 0.0; n = 0; flips = 0; r_all = []
        for c in range(a["sel"].shape[0]):
            for i in range(a["sel"].shape[1]):
                T = int(a["nt"][i])
                if T == 0: continue
                o = int(a["off"][i])
                x, y = a["te"][c, o:o+T], b["te"][c, o:o+T]
                fin = np.isfinite(x) & np.isfinite(y)
                assert np.array_equal(np.isinf(x), np.isinf(y))
                if fin.any():
                    r = np.abs(x[fin] - y[fin]) / np.maximum(1e-3, 1e-5 * np.abs(x[fin]))
                    r_all.append(r); n += fin.sum()
                if a["sel"][c, i] != b["sel"][c, i]:
                    flips += 1; break
        r = np.concatenate(r_all)
        print("%s vs %s: %d trials, difference / tolerance: max %.4f  p99 %.4f  median %.5f ; selection flips %d" % (
            lib.name, libs[0].name, n, r.max(), np.quantile(r, 0.99), np.median(r), flips), flush=True)
