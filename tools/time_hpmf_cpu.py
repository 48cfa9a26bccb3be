"""CPU time of the reference's own calc_sasa / init_hpmf_calculator closures (build container only: needs
/root/reference) and of the NumPy oracle, per structure, on the 76-residue fixture structure and on four copies of it
placed 200 A apart (4888 atoms - the size of a config-3 structure).  One core, NumPy.  Prints one JSON line."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tests" / "golden"))

from golden_utils import GOLDEN  # noqa: E402
from mcsce_b200.hpmf import load_sasa_tables  # noqa: E402
from oracle import hpmf_numpy as O  # noqa: E402

z = np.load(GOLDEN / "hpmf.npz", allow_pickle=False)
g = {k[len("helix76_"):]: z[k] for k in z.files if k.startswith("helix76_")}
types = [str(t) for t in g["types"]]
f = g["atom_filter"]
n = int(f.sum())
bonded = np.zeros((n, n), bool); bonded[g["bond_i"], g["bond_j"]] = True; bonded |= bonded.T
x = g["coords"][0]


def tile(k):
    xs = np.concatenate([x + np.float32(200.0 * i) for i in range(k)])
    ff = np.concatenate([f] * k)
    tt = types * k
    bb = np.zeros((n * k, n * k), bool)
    for i in range(k):
        bb[i * n:(i + 1) * n, i * n:(i + 1) * n] = bonded
    return xs, ff, tt, bb


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


out = {"cpu": "build container, one core"}
tb = load_sasa_tables()
try:
    from ref_loader import load_reference
    load_reference(0)
    from mcsce.libs import libhpmf
    have_ref = True
except Exception as e:                      # the reference is not on the GPU box
    have_ref = False
    out["reference"] = "unavailable: %s" % e
for k in (1, 4):
    xs, ff, tt, bb = tile(k)
    key = "atoms_%d" % len(xs)
    out[key] = {"typed": int(ff.sum()), "oracle_s": best(lambda: O.hpmf(xs, ff, tt, bb, tb))}
    if have_ref:
        fn = libhpmf.init_hpmf_calculator(tt, ff, bb.astype(float))
        v_ref = fn(xs)
        out[key]["reference_s"] = best(lambda: fn(xs))
        out[key]["v_reference"] = float(v_ref)
        out[key]["v_oracle"] = float(O.hpmf(xs, ff, tt, bb, tb)[0])
print(json.dumps(out))
