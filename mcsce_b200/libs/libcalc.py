"""Drop-in for ``mcsce.libs.libcalc.place_sidechain_template`` (libcalc.py:419-482): the two
alignment quaternions are host tables (mcsce_b200.libs.placement), their application to the template
atoms runs on the device (``mcsce_place_template``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from mcsce_b200 import _lib
from mcsce_b200.libs.placement import placement_quaternions


def place_sidechain_template(bb_cnf, ss_template, device=None):
    """float64 (M, 3): template ``ss_template`` (M, 3; CA at the origin, atoms 0/2 = N/C) placed on the
    backbone ``bb_cnf`` (3, 3) N, CA, C.  All template atoms are returned."""
    from mcsce_b200.libs.libenergy import _device

    bb = np.asarray(bb_cnf, dtype=np.float32)
    tmpl = np.ascontiguousarray(ss_template, dtype=np.float32)
    q1, q2, ca = placement_quaternions(bb, tmpl[0], tmpl[2])
    out = np.empty((len(tmpl), 3), dtype=np.float64)
    lib = _lib.load_library()
    p = lambda a, ty: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(ty))  # noqa: E731
    q1 = np.ascontiguousarray(q1, np.float64); q2 = np.ascontiguousarray(q2, np.float64); ca = np.ascontiguousarray(ca, np.float32)
    _lib.check(lib.mcsce_place_template(_device() if device is None else int(device), len(tmpl), p(tmpl, C.c_float),
                                        p(q1, C.c_double), p(q2, C.c_double), p(ca, C.c_float), p(out, C.c_double)))
    return out
