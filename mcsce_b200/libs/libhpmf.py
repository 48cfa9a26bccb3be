"""Drop-in for ``mcsce.libs.libhpmf`` (libhpmf.py:17-111): hydrophobic potential of mean force
(Lin, Fawzi, Head-Gordon 2007) over exposed carbon pairs, evaluated by the CUDA kernels
``sasa_kernel`` + ``hpmf_pair_kernel`` through ``mcsce_hpmf_energy`` (include/mcsce_b200.h).

Same names and closure contract as the reference; the closure also takes a batch of structures
(B, N, 3) and then returns (B,) values - the form the growth results come in."""
from __future__ import annotations

import numpy as np

from mcsce_b200.hpmf import HpmfCalculator, convert_atom_types, load_sasa_tables  # noqa: F401


def __getattr__(name):
    tb = load_sasa_tables()["hpmf"]
    table = {"C1": ("C", 0), "C2": ("C", 1), "C3": ("C", 2), "W1": ("W", 0), "W2": ("W", 1), "W3": ("W", 2),
             "H1": ("H", 0), "H2": ("H", 1), "H3": ("H", 2)}
    if name == "AC":
        return tb["AC"]
    if name in table:
        k, i = table[name]
        return tb[k][i]
    raise AttributeError(name)


def _device():
    import os
    return int(os.environ.get("LOCAL_RANK", os.environ.get("MCSCE_B200_DEVICE", "0")))


def init_hpmf_calculator(sasa_atom_types, atom_filter, bond_connectivities, threshold=None):
    """``init_hpmf_calculator(sasa_atom_types, atom_filter, bond_connectivities, threshold=AC)`` ->
    ``calculate(coords)``; ``coords`` covers all atoms (``atom_filter`` picks the typed ones).
    (N, 3) -> float as in the reference (libhpmf.py:92-109); (B, N, 3) -> (B,) float64 ndarray."""
    calc_dev = HpmfCalculator(list(sasa_atom_types), atom_filter, bond_connectivities, threshold=threshold,
                              device=_device())

    def calculate(coords):
        coords = np.asarray(coords, dtype=np.float32)
        v = calc_dev.energy(coords).cpu().numpy()
        return float(v) if coords.ndim == 2 else v

    calculate.calculator = calc_dev
    return calculate
