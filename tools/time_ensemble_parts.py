"""Where the time of create_side_chain_ensemble (cfg3, N conformers) goes: growth, PDB files, the rest."""
import sys, time, tempfile, os
from functools import partial
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from mcsce_b200 import synth
from mcsce_b200.structure import Structure
from mcsce_b200.core import side_chain_builder as scb
from mcsce_b200.libs.libenergy import prepare_energy_function
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
s = Structure(synth.config_backbone(3)).build().remove_side_chains([])
scb.initialize_func_calc(partial(prepare_energy_function, batch_size=16, forcefield=None, terms=["lj", "clash", "coulomb"]), structure=s)
d = tempfile.mkdtemp()
print("output dir", d, "cores", os.cpu_count())
scb.create_side_chain_ensemble(s, 64, 300, d)
t0 = time.perf_counter(); coords, energy, ok = scb._grow(s.coords, 1 / (300 * 0.008314462), n); t1 = time.perf_counter()
w = scb._write_pdbs(coords, ok, d); t2 = time.perf_counter()
size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".pdb"))
print("grow %.2f s | write %d PDB files %.2f s (%.0f MB -> %.0f MB/s)" % (t1 - t0, w, t2 - t1, size / 1e6, size / 1e6 / (t2 - t1)))
t3 = time.perf_counter(); confs, okl = scb.create_side_chain_ensemble(s, n, 300, d); t4 = time.perf_counter()
print("create_side_chain_ensemble %.2f s -> %.0f conformers/s" % (t4 - t3, n / (t4 - t3)))
