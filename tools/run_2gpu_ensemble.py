"""torchrun check of the sharded drop-in call: ranks grow contiguous conformer ranges, rank 0 gathers (NCCL),
writes the files and compares with growing all conformers on its own GPU."""
import os, sys, tempfile
from functools import partial
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from golden_utils import GoldenCase
from mcsce_b200.core import side_chain_builder as scb
from mcsce_b200.libs.libenergy import prepare_energy_function
case = GoldenCase("pep12")
s = case.structure()
scb.set_seed(17)
scb.initialize_func_calc(partial(prepare_energy_function, batch_size=16, forcefield=None, terms=["lj", "clash", "coulomb"]), structure=s)
d = tempfile.mkdtemp() if dist.get_rank() == 0 else "/nonexistent"
n = 11
confs, ok = scb.create_side_chain_ensemble(s, n, 300, d)
if dist.get_rank() == 0:
    eng = scb._state["engine"]
    ref = eng.grow(n, 1 / (300 * 0.008314462), seed=17, conf_id0=0, host_out=True)
    lines = open(d + "/energies.csv").read().strip().splitlines()[1:]
    got = {int(os.path.basename(l.split(",")[0]).split(".")[0]): float(l.split(",")[1]) for l in lines}
    assert len(ok) == n and sum(ok) == int(ref["ok"].sum()) == len(got), (ok, ref["ok"])
    for c in range(n):
        assert bool(ok[c]) == bool(ref["ok"][c])
        if ok[c]:
            assert abs(got[c] - ref["energy"][c]) < 1e-5, (c, got[c], ref["energy"][c])
            assert np.abs(confs[c].coords - np.round(ref["coords"][c], 3)[np.argsort(scb._state["system"].resnums, kind="mergesort")]).max() < 2e-3
    print("2-GPU sharded ensemble == single-GPU growth: OK (%d conformers, %d ranks)" % (n, dist.get_world_size()))
else:
    assert confs == [] and ok == []
dist.barrier()
dist.destroy_process_group()
