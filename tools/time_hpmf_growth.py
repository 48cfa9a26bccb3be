"""Wall time of the step-by-step growth with and without the hpmf term:
    python tools/time_hpmf_growth.py [cfg] [n_conf]
Prints one JSON line."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

from golden_utils import library  # noqa: E402
from mcsce_b200 import synth  # noqa: E402
from mcsce_b200.engine import GrowthEngine  # noqa: E402
from mcsce_b200.hpmf import growth_term  # noqa: E402
from mcsce_b200.structure import Structure  # noqa: E402
from mcsce_b200.system import build_system  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_conf = int(sys.argv[2]) if len(sys.argv) > 2 else 32
beta = 1.0 / (300 * 0.008314462)
lib, ptm = library()
s = Structure(synth.config_backbone(cfg)).build().remove_side_chains([])
sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
term = growth_term(eng)


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


t_fused, a = timed(lambda: eng.grow(n_conf, beta, seed=1))
t_step, b = timed(lambda: eng.grow_stepwise(n_conf, beta, seed=1))
t_hpmf, c = timed(lambda: eng.grow_stepwise(n_conf, beta, seed=1, extra_term=term), reps=1)
t_bf = None
if "--brute" in sys.argv:
    bf = growth_term(eng, term.calculator, brute_force=True)
    t_bf, _ = timed(lambda: eng.grow_stepwise(n_conf, beta, seed=1, extra_term=bf), reps=1)
n_trials = int(eng.n_trials_total)
print(json.dumps({"workload": "cfg%d" % cfg, "conformers": n_conf, "atoms": sysm.n_atoms, "trials_per_conformer": n_trials,
                  "fused_s": t_fused, "stepwise_s": t_step, "stepwise_hpmf_s": t_hpmf, "stepwise_hpmf_brute_force_s": t_bf,
                  "hpmf_trials_per_s": n_conf * n_trials / max(1e-9, t_hpmf - t_step),
                  "ok": [float(x["ok"].float().mean()) for x in (a, b, c)],
                  "mean_energy": [float(x["energy"][x["ok"].bool()].mean()) for x in (a, b, c)]}))
