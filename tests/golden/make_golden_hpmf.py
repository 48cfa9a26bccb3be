"""Golden fixtures and constant tables for the hydrophobic-PMF term.

    python tests/golden/make_golden_hpmf.py          (build container only)

The reference cannot run this term through ``prepare_energy_function`` (``connectivity_matrix``
is commented out at ``libs/libenergy.py:110-113`` and used at ``:217``; the closure is then called
with two arguments, ``:807``, but takes one, ``libs/libhpmf.py:92``).  Its two building blocks
*do* run when called directly, and they are what this script pins:

* ``libs/libsasa.py:76-109``  ``calc_sasa(atom_types, bond_connectivities)(coords)``
* ``libs/libhpmf.py:22-111``  ``convert_atom_types`` and ``init_hpmf_calculator(...)(coords)``

The bond-connectivity matrix is built by the reference's own mask routine exactly as the
commented-out lines intend (``bonds_1_mask`` mirrored into an N x N matrix), restricted to the atoms
that carry a SASA type.  Inputs are the final structures the reference grew for the growth fixtures
(``tests/golden/*.npz``), each also jittered and compressed so that some carbons fall under the
6 A^2 exposure threshold.

Also written: ``mcsce_b200/data/sasa_tables.json.gz`` - the Hasel-Hendrickson-Still P/R parameters
(``libsasa.py:21-74``), the residue/atom -> SASA type map (``core/definitions.py:181-511``) and the
three-Gaussian constants (``libhpmf.py:17-20``), read from the imported reference modules, because
the reference tree does not exist on the GPU box.
"""
import gzip
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parents[1]))

from ref_loader import load_reference  # noqa: E402

CASES = ["pep12", "std20", "bundle40", "helix76"]


def write_tables(out_path):
    load_reference(0)
    from mcsce.core.definitions import aa_atom_type_mappings
    from mcsce.libs import libhpmf, libsasa

    packed = {
        "format": 1,
        "source": "THGLab/MCSCE libs/libsasa.py, libs/libhpmf.py, core/definitions.py (values read from the imported modules)",
        "P": libsasa.P_PARAMS, "R": libsasa.R_PARAMS,
        "pij_bonded": libsasa.PIJ_BONDED, "pij_non_bonded": libsasa.PIJ_NON_BONDED,
        "r_solvent": 1.4,
        "atom_types": {res: {a: t for a, t in atoms.items()} for res, atoms in aa_atom_type_mappings.items()},
        "hpmf": {"C": [libhpmf.C1, libhpmf.C2, libhpmf.C3], "W": [libhpmf.W1, libhpmf.W2, libhpmf.W3],
                 "H": [libhpmf.H1, libhpmf.H2, libhpmf.H3], "AC": libhpmf.AC},
    }
    with gzip.GzipFile(out_path, "wb", mtime=0) as f:
        f.write(json.dumps(packed, sort_keys=True, separators=(",", ":")).encode())
    return out_path


def variants(coords, seed):
    """The grown structure, two jittered copies and two copies compressed towards the centroid."""
    rng = np.random.default_rng(seed)
    c = coords.astype(np.float32)
    centre = c.mean(0, keepdims=True)
    out = [c,
           (c + rng.normal(0, 0.15, c.shape)).astype(np.float32),
           (c + rng.normal(0, 0.4, c.shape)).astype(np.float32),
           (centre + (c - centre) * np.float32(0.85)).astype(np.float32),
           (centre + (c - centre) * np.float32(0.7) + rng.normal(0, 0.1, c.shape)).astype(np.float32)]
    return np.stack(out)


def run():
    from golden_utils import GoldenCase

    load_reference(0)
    from mcsce.core.build_definitions import bonds_equal_1_inter, forcefields
    from mcsce.libs import libhpmf, libsasa
    from mcsce.libs.libenergy import create_bonds_apart_mask_for_ij_pairs, create_residue_data_dict

    ff = forcefields["Amberff14SB"](Cterminal="OXT", Nterminal="HN")
    out = {}
    for k, name in enumerate(CASES):
        case = GoldenCase(name)
        conf = case.conf(0)
        names = [str(n) for n in conf["final_names"]]
        resnums = [int(r) for r in conf["final_resnums"]]
        s = case.structure()
        label_of = {int(n): str(l) for n, l in zip(s.res_nums, s.res_labels)}
        reslabels = [label_of[r] for r in resnums]
        N = len(names)
        # the connectivity the commented-out lines of libenergy.py:110-113 would have produced
        residue_data = create_residue_data_dict(names, resnums, reslabels)
        bonds_1 = create_bonds_apart_mask_for_ij_pairs(residue_data, N, ff.res_topology, bonds_equal_1_inter,
                                                       base_bool=False)
        conn = np.zeros((N, N))
        conn[np.triu_indices(N, k=1)] = bonds_1
        conn += conn.T
        atom_filter, types = libhpmf.convert_atom_types(names, resnums, reslabels)
        conn_typed = conn[np.ix_(atom_filter, atom_filter)]
        sasa_fn = libsasa.calc_sasa(types, conn_typed)
        hpmf_fn = libhpmf.init_hpmf_calculator(types, atom_filter, conn_typed)
        xs = variants(conf["final_coords"], 100 + k)
        sasa = np.stack([sasa_fn(x[atom_filter]) for x in xs])
        v = np.array([hpmf_fn(x) for x in xs])
        bi, bj = np.nonzero(np.triu(conn_typed))
        carbon = np.array([t.startswith("C") for t in types])
        print("[%s] atoms %d typed %d carbons %d  V %s  carbons under threshold %s" % (
            name, N, atom_filter.sum(), carbon.sum(), np.round(v, 4), [(sa[carbon] <= 6).sum() for sa in sasa]))
        pre = name + "_"
        out[pre + "names"] = np.array(names)
        out[pre + "resnums"] = np.array(resnums, dtype=np.int32)
        out[pre + "reslabels"] = np.array(reslabels)
        out[pre + "atom_filter"] = atom_filter
        out[pre + "types"] = np.array(types)
        out[pre + "bond_i"] = bi.astype(np.int32)
        out[pre + "bond_j"] = bj.astype(np.int32)
        out[pre + "coords"] = xs
        out[pre + "sasa"] = sasa
        out[pre + "v"] = v
    out["cases"] = np.array(CASES)
    np.savez_compressed(HERE / "hpmf.npz", **out)
    print("wrote", HERE / "hpmf.npz")


if __name__ == "__main__":
    print(write_tables(HERE.parents[1] / "mcsce_b200" / "data" / "sasa_tables.json.gz"))
    run()
