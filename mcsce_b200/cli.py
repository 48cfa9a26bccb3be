"""Drop-in for the ``mcsce`` command line (mcsce/cli.py:12-223): same arguments, same outputs
(``<out>/<n>.pdb``, ``energies.csv``, ``log.csv``).  ``-w/--n_worker`` sized the reference's process
pool and is accepted and ignored (one process per GPU; launch under ``torchrun`` to use several GPUs);
``-b/--batch_size`` only bounded the reference's distance scratch and is ignored as well."""
import argparse
import os
from datetime import datetime
from functools import partial

import numpy as np

parser = argparse.ArgumentParser()
parser.add_argument("input_structure", help="backbone PDB file, or a directory whose *.pdb files are all processed")
parser.add_argument("n_conf", type=int, help="conformers to attempt per input backbone (failed growths count)")
parser.add_argument("-w", "--n_worker", type=int, default=None, help="Accepted for compatibility (the reference's process-pool size); the GPU path uses one process per GPU")
parser.add_argument("-o", "--output_dir", default=None, help="where results go; default: <input name>_mcsce next to the working directory")
parser.add_argument("-s", "--same_structure", action="store_true", default=False, help="All structures in the folder share one amino acid sequence: prepare the tables once")
parser.add_argument("-b", "--batch_size", default=16, type=int, help="Accepted for compatibility; not used by the GPU path")
parser.add_argument("-m", "--mode", choices=["ensemble", "simple", "exhaustive"], default="ensemble", help="ensemble: write every conformer; simple: first clash-free conformer; exhaustive: lowest-energy conformer of n_conf")
parser.add_argument("-f", "--fix", default=None, help="residue ids whose side chains are retained, e.g. 2-5+10-13 (same-structure mode only)")
parser.add_argument("-l", "--logfile", default="log.csv", help="CSV log: file, number grown, wall time")
parser.add_argument("-v", "--verbose", default=False, help="print the full missing-atom / HIS renaming warnings")
parser.add_argument("--seed", type=int, default=0, help="seed of the on-device random streams (extension)")
parser.add_argument("--hpmf", action="store_true", default=False, help="add the hydrophobic PMF term to lj + clash + coulomb (extension; the reference's CLI fixes the three terms, cli.py:149-151)")


def parse_fix(fix):
    """'2-5+10-13' -> [2,3,4,5,10,11,12,13]; None on a syntax error (cli.py:118-141)."""
    fix = fix.strip()
    out = []
    for chunk in fix.split("+"):
        if "-" in chunk:
            parts = chunk.split("-")
            if len(parts) != 2:
                return None
            out += list(range(int(parts[0]), int(parts[1]) + 1))
        else:
            out.append(int(chunk))
    return out


def main(input_structure, n_conf, n_worker, output_dir, logfile, mode, fix, batch_size=4, same_structure=False,
         verbose=False, seed=0, hpmf=False):
    from mcsce_b200 import sharding
    from mcsce_b200.core import side_chain_builder as scb
    from mcsce_b200.libs.libenergy import prepare_energy_function
    from mcsce_b200.structure import Structure, read_structure_and_check

    root = sharding.world()[0] == 0
    base_dir = output_dir
    if root:
        with open(logfile, "w") as f:
            f.write("PDB name,Succeeded,Time used\n")
    if isinstance(fix, str):
        if fix.find("-") + fix.find("+") <= -2 and not fix.strip().isdigit():
            print("--fix argument syntax error. Abort")
            return
        elif not same_structure:
            print("--fix argument ignored")
    if input_structure[-3:].upper() == "PDB":
        all_pdbs = [input_structure]
    else:
        folder = input_structure if input_structure.endswith("/") else input_structure + "/"
        all_pdbs = sorted(folder + f for f in os.listdir(folder) if f[-3:].upper() == "PDB")
    terms = ["lj", "clash", "coulomb"] + (["hpmf"] if hpmf else [])
    creator = partial(prepare_energy_function, batch_size=batch_size, forcefield=None, terms=terms)
    scb.set_seed(seed)
    fix_idxs = []
    if same_structure:
        s = Structure(all_pdbs[0]).build()
        if fix is not None:
            fix_idxs = parse_fix(fix)
            if fix_idxs is None:
                print("--fix argument syntax error. Abort")
                return
            if np.any(np.array(fix_idxs) > s.res_nums[-1]):
                print("--fix residue id out of range")
                return
        s = read_structure_and_check(all_pdbs[0], fix_idxs, verbose)
        scb.initialize_func_calc(creator, structure=s, retain_idxs=fix_idxs)
    for f in all_pdbs:
        print("Now working on", f)
        t0 = datetime.now()
        stem = os.path.splitext(os.path.basename(f))[0]
        out_dir = (stem + "_mcsce") if base_dir is None else (base_dir + "/" + stem + "_mcsce")
        if root and mode == "ensemble":
            os.makedirs(out_dir, exist_ok=True)
        s = read_structure_and_check(f, fix_idxs, verbose=verbose)
        if not same_structure:
            scb.initialize_func_calc(creator, structure=s)
        if mode == "ensemble":
            _, success = scb.create_side_chain_ensemble(s, n_conf, temperature=300, retain_resi=fix_idxs, save_path=out_dir,
                                                        parallel_worker=n_worker)
            if root:
                with open(logfile, "a") as log:
                    log.write("%s,%d,%s\n" % (f, sum(success), str(datetime.now() - t0)))
        else:
            best = scb.create_side_chain(s, n_conf, 300, parallel_worker=n_worker, retain_resi=fix_idxs,
                                         return_first_valid=(mode == "simple"))
            if root:
                if best is not None:
                    best.write_PDB(out_dir + ".pdb")
                with open(logfile, "a") as log:
                    log.write("%s,%d,%s\n" % (f, 1 if best is not None else 0, str(datetime.now() - t0)))
    print("All finished!")


def maincli():
    args = vars(parser.parse_args())
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1:                                   # launched with torchrun: one process per GPU
        import torch
        import torch.distributed as dist

        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        main(**args)
    finally:
        if ws > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    maincli()
