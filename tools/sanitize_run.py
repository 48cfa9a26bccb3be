"""Small run for compute-sanitizer: golden pep12 + ptm14 growths (replay + RNG), K1/K2/K3 API calls."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from test_gpu_parity import engine_for, oracle_env_sequence
for name in ("pep12", "ptm14"):
    case, s, sysm, eng = engine_for(name)
    out = eng.grow(3, case.beta, seed=1, dump=True, no_graph=True)
    out = eng.grow(3, case.beta, seed=1, env_split=3)
    out = eng.grow(40, case.beta, seed=2, host_out=True)
    envs = oracle_env_sequence(case, 0, s, sysm)
    for i, g in list(case.steps_by_level(0).items())[:3]:
        eng.place_trials(i, g["chis"])
        eng.score_trials(i, g["trial"][None], envs[i][None])
    torch.cuda.synchronize()
print("sanitize run done")
