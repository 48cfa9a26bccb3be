"""Drop-in for ``mcsce.libs.libscbuild.rotate_sidechain`` (libscbuild.py:56-206): set chi_1..chi_n
on a residue template.  The rotation program (dihedral atoms, axes, moving sets) is data in
``mcsce_b200.chem.tables``; the rotations run on the device (``mcsce_rotate_template``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from mcsce_b200 import _lib
from mcsce_b200.chem.tables import load_tables


def rotate_sidechain(res_type, tors, device=None):
    """(coords (n_tmpl, 3) float32, sidechain_idx) with the template's chi angles set to ``tors`` degrees."""
    assert res_type not in ("ALA", "GLY", "PCA")
    from mcsce_b200.libs.libenergy import _device

    out, sc_idx = rotate_sidechain_batch(res_type, np.asarray(tors, dtype=np.float64)[None], _device() if device is None else device)
    return out[0], sc_idx


def rotate_sidechain_batch(res_type, tors, device=0):
    """Batched form: ``tors`` (rows, n_chi) degrees -> (rows, n_tmpl, 3) float32."""
    t = load_tables()
    tm = t.templates[res_type]
    tors = np.asarray(tors, dtype=np.float64)
    rows = np.zeros((len(tors), _lib.MAX_CHI))
    rows[:, :min(tors.shape[1], _lib.MAX_CHI)] = tors[:, :_lib.MAX_CHI]
    # rotations beyond the residue's real chi angles (template positions reused by the reference's generic
    # rules) are not applied: see ResidueTemplate._count_sound
    n_rot = min(tors.shape[1], tm.n_sound, _lib.MAX_CHI_ROT)
    axis = np.zeros((_lib.MAX_CHI_ROT, 2), np.int32); move = np.zeros(_lib.MAX_CHI_ROT, np.uint32); ori = np.zeros(_lib.MAX_CHI_ROT)
    for k, chi in enumerate(tm.chis[:n_rot]):
        axis[k] = chi["axis"]; move[k] = sum(1 << m for m in chi["moving"]); ori[k] = chi["ori"]
    xyz = np.ascontiguousarray(tm.coords, np.float32)
    out = np.empty((len(rows), tm.n_atoms, 3), np.float32)
    lib = _lib.load_library()
    p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))  # noqa: E731
    _lib.check(lib.mcsce_rotate_template(int(device), tm.n_atoms, p(xyz, C.c_float), n_rot, p(axis, C.c_int32),
                                         p(move, C.c_uint32), p(ori, C.c_double), len(rows), p(rows, C.c_double),
                                         p(out, C.c_float)))
    return out, tm.sc_idx.copy()
