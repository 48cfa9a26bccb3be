"""End-to-end drop-in call: create_side_chain_ensemble on cfg3 with N conformers (PDB files included)."""
import sys, time, tempfile
from functools import partial
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mcsce_b200 import synth
from mcsce_b200.structure import Structure
from mcsce_b200.core import side_chain_builder as scb
from mcsce_b200.libs.libenergy import prepare_energy_function
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
s = Structure(synth.config_backbone(3)).build().remove_side_chains([])
t0 = time.perf_counter()
scb.initialize_func_calc(partial(prepare_energy_function, batch_size=16, forcefield=None, terms=["lj", "clash", "coulomb"]), structure=s)
t1 = time.perf_counter()
d = tempfile.mkdtemp()
confs, ok = scb.create_side_chain_ensemble(s, 64, 300, d)        # warm-up
t2 = time.perf_counter()
confs, ok = scb.create_side_chain_ensemble(s, n, 300, d)
t3 = time.perf_counter()
print("initialize_func_calc %.2f s | ensemble of %d: %.2f s -> %.0f conformers/s incl. %d PDB files + energies.csv + Structure objects"
      % (t1 - t0, n, t3 - t2, n / (t3 - t2), sum(ok)))
