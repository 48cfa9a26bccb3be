"""Fused hpmf growth with and without the residue lists: where do they differ?"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("MCSCE_B200_ALLOW_SYNTHETIC", "1")
import numpy as np, bench
from mcsce_b200.engine import GrowthEngine
from mcsce_b200.hpmf import calculator_for_labels
s, sysm, n_conf, mode, desc = bench.build_workload("cfg3")
eng = GrowthEngine(0).set_system(sysm).set_backbone(s.coords)
calc = calculator_for_labels([str(a) for a in sysm.names], [int(r) for r in sysm.resnums], [str(r) for r in sysm.resnames])
u = np.random.default_rng(8).random((4, sysm.n_res))
a = eng.grow(4, bench.BETA, seed=21, uniforms=u, dump=True, hpmf=calc)
ta = a["trial_energy"].cpu().numpy(); chis = a["chis"].cpu().numpy()
os.environ["MCSCE_B200_HPMF_NO_LISTS"] = "1"
import subprocess
# second process: the knob is read once per process
if len(sys.argv) > 1:
    np.save("/tmp/ta_nolist.npy", ta); sys.exit(0)
subprocess.run([sys.executable, __file__, "child"], check=True, env=dict(os.environ))
tb = np.load("/tmp/ta_nolist.npy")
fin = np.isfinite(ta) & np.isfinite(tb)
d = np.abs(ta - tb); d[~fin] = 0
print("max |dE| with vs without lists:", d.max())
for c in range(4):
    for st in sysm.steps:
        if st.kind != "grow": continue
        o = eng.trial_off[st.index]; x = d[c, o:o + len(st.prob)].max()
        if x > 1e-7: print("conf", c, "residue", st.index, st.resname, "n_chi", st.n_chi, "max diff", x)
