"""Time the UNMODIFIED reference (``baseline/_ref``, installed by ``baseline/install_reference.py``) on the host cores.

    python baseline/stock_reference.py --workload cfg3 --conformers 16 --workers 8

Follows the reference's own CLI (``mcsce/cli.py:149-151, 197-205``): ``initialize_func_calc(partial(prepare_energy_function,
batch_size, forcefield=Amberff14SB, terms=[lj, clash, coulomb]), structure=s)``, then
``create_side_chain_ensemble(s, n_conf, 300, out_dir, parallel_worker=W)`` - a ``multiprocessing.Pool`` over conformers.
Setup and ensemble are timed separately (the setup is not the hot path).  Inputs: the synthetic backbones of
``mcsce_b200.synth`` and the synthetic rotamer library packed into the install.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import tempfile
import time
from functools import partial
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--conformers", type=int, default=0, help="default: 2 per worker")
    ap.add_argument("--workers", type=int, default=0, help="default: os.cpu_count()")
    ap.add_argument("--batch-size", type=int, default=16)
    args = ap.parse_args()
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "mcsce" / "core" / "side_chain_builder.py").exists():
        print(json.dumps({"impl": "stock_reference", "unavailable": "baseline/_ref is not installed (python baseline/install_reference.py)"}))
        return
    sys.path.insert(0, str(ROOT))
    from mcsce_b200 import synth                      # synthetic backbone text only

    cfg = int(args.workload.replace("cfg", ""))
    pdb_text = synth.config_backbone(cfg)
    sys.path.insert(0, str(ref))
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "numba_cache_stock_ref"))
    t0 = time.perf_counter()
    from mcsce.core import side_chain_builder as scb
    from mcsce.core.build_definitions import forcefields
    from mcsce.libs.libenergy import prepare_energy_function
    from mcsce.libs.libstructure import Structure
    import_s = time.perf_counter() - t0

    W = args.workers or (os.cpu_count() or 1)
    n = args.conformers or 2 * W
    tmp = tempfile.mkdtemp(prefix="stock_ref_")
    pdb = os.path.join(tmp, "backbone.pdb")
    Path(pdb).write_text(pdb_text)
    s = Structure(pdb)
    s.build()
    s = s.remove_side_chains()
    ff = forcefields["Amberff14SB"](Cterminal="OXT", Nterminal="HN")
    t0 = time.perf_counter()
    scb.initialize_func_calc(partial(prepare_energy_function, batch_size=args.batch_size, forcefield=ff,
                                     terms=["lj", "clash", "coulomb"]), structure=s)
    init_s = time.perf_counter() - t0
    out_dir = os.path.join(tmp, "out")
    os.makedirs(out_dir)
    t0 = time.perf_counter()
    confs, ok = scb.create_side_chain_ensemble(s, n, 300, out_dir, parallel_worker=W)
    ens_s = time.perf_counter() - t0
    print(json.dumps({
        "impl": "stock_reference", "workload": args.workload, "residues": len(s.residue_types), "conformers": n, "workers": W,
        "host_cores": os.cpu_count(), "import_seconds": import_s, "initialize_func_calc_seconds": init_s,
        "ensemble_seconds": ens_s, "conformers_per_second": n / ens_s, "succeeded": int(sum(bool(x) for x in ok)),
        "entry": "create_side_chain_ensemble(structure, n, 300, out_dir, parallel_worker=W) (side_chain_builder.py:284-352), "
                 "PDB files written like the CLI does",
        "note": "unmodified reference from baseline/_ref; synthetic rotamer library and backbone"}))


if __name__ == "__main__":
    main()
