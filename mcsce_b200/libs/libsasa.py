"""Drop-in for ``mcsce.libs.libsasa`` (libsasa.py:76-109): Hasel-Hendrickson-Still approximate
solvent-accessible surface areas, evaluated by the CUDA kernel ``sasa_kernel`` through
``mcsce_sasa`` (include/mcsce_b200.h).  Same closure contract; the closure also takes a batch."""
from __future__ import annotations

import numpy as np

from mcsce_b200.hpmf import HpmfCalculator, load_sasa_tables


def __getattr__(name):
    # P_PARAMS / R_PARAMS / PIJ_* as module attributes, like the reference module (libsasa.py:21-74)
    tb = load_sasa_tables()
    key = {"P_PARAMS": "P", "R_PARAMS": "R", "PIJ_BONDED": "pij_bonded", "PIJ_NON_BONDED": "pij_non_bonded"}.get(name)
    if key is None:
        raise AttributeError(name)
    return tb[key]


def _device():
    import os
    return int(os.environ.get("LOCAL_RANK", os.environ.get("MCSCE_B200_DEVICE", "0")))


def calc_sasa(atom_types, bond_connectivities, r_solvent=1.4):
    """``calc_sasa(atom_types, bond_connectivities, r_solvent)`` -> ``calc(coords)``.

    ``coords`` (N, 3) -> (N,) float64 ndarray as in the reference; (B, N, 3) -> (B, N)."""
    n = len(atom_types)
    calc_dev = HpmfCalculator(list(atom_types), np.ones(n, dtype=bool), bond_connectivities, r_solvent=r_solvent,
                              device=_device())

    def calc(coords):
        return calc_dev.sasa(np.asarray(coords, dtype=np.float32).reshape((-1, n, 3) if np.ndim(coords) == 3
                                                                          else (n, 3))).cpu().numpy()

    calc.calculator = calc_dev
    return calc
