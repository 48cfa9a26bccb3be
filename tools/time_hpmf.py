"""Device time of the A7 kernels on grown conformers: python tools/time_hpmf.py [cfg] [n_conf]
Prints one JSON line (SASA kernel ms, carbon-pair kernel ms, pair tests per second)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from golden_utils import library  # noqa: E402
from mcsce_b200 import synth  # noqa: E402
from mcsce_b200.engine import GrowthEngine  # noqa: E402
from mcsce_b200.hpmf import calculator_for_labels  # noqa: E402
from mcsce_b200.structure import Structure  # noqa: E402
from mcsce_b200.system import build_system  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_conf = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
lib, ptm = library()
s = Structure(synth.config_backbone(cfg)).build().remove_side_chains([])
sysm = build_system(s, rotamer_library=lib, ptm_library=ptm)
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
out = eng.grow(n_conf, 1.0 / (300 * 0.008314462), seed=1)
coords = out["coords"][out["ok"].bool()].contiguous()
calc = calculator_for_labels([str(a) for a in sysm.names], [int(r) for r in sysm.resnums], [str(r) for r in sysm.resnames])
ms = []
for _ in range(5):
    v = calc.energy(coords)
    torch.cuda.synchronize()
    ms.append(calc.last_ms())
sasa_ms, pair_ms = np.median([m[0] for m in ms]), np.median([m[1] for m in ms])
B = coords.shape[0]
n_c = int(calc.carbon.sum())
print(json.dumps({"workload": "cfg%d" % cfg, "structures": B, "atoms": sysm.n_atoms, "typed": calc.n_typed, "carbons": n_c,
                  "sasa_ms": sasa_ms, "pair_ms": pair_ms,
                  "sasa_pair_tests_per_s": B * calc.n_typed ** 2 / (sasa_ms * 1e-3),
                  "carbon_pair_tests_per_s": B * n_c ** 2 / (pair_ms * 1e-3),
                  "structures_per_s": B / ((sasa_ms + pair_ms) * 1e-3),
                  "v_mean": float(v.mean()), "v_min": float(v.min()), "v_max": float(v.max())}))
