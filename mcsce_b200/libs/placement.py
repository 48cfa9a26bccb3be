"""Per-residue rigid placement of a residue template onto the backbone (host tables).

``place_sidechain_template`` (``mcsce/libs/libcalc.py:419-482``) aligns the template's N-CA
vector with the backbone's and then the N-CA-C plane normals.  Both rotations depend only on
the backbone N, CA, C and on template atoms 0 (N) and 2 (C), which no chi rotation moves, so
they are identical for every trial and every conformer of a residue: they are computed here
once per backbone as two quaternions (O(L) host work) and applied on the GPU by the
placement kernel.  float32/float64 mix as in the reference: vectors, cross products, norms
and dot products float32 (numba lowers ``np.linalg.norm``/``np.dot`` on float32 to BLAS
snrm2/sdot), clamp and arccos float64, quaternion components float64.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import blas as _blas

_F32, _F64 = np.float32, np.float64


def _unit32(v):
    v = np.ascontiguousarray(v, dtype=_F32)
    return v / _F32(_blas.snrm2(v))


def _angle(v1, v2):
    d = _F64(_F32(_blas.sdot(_unit32(v1), _unit32(v2))))
    return np.arccos(min(max(d, _F64(-1.0)), _F64(1.0)))


def _cross32(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dtype=_F32)


def _quat(axis_unit32, angle):
    s = np.sin(_F64(angle) / 2)
    a = axis_unit32.astype(_F64)
    return np.array([np.cos(_F64(angle) / 2), s * a[0], s * a[1], s * a[2]], dtype=_F64)


def _rotate(q, p):
    """q * (0,p) * conj(q), term order of libcalc.py:359-366."""
    b1, b2, b3, b4 = q
    x, y, z = (_F64(v) for v in p)
    zero = _F64(0.0)
    c1 = (b1 * zero) - (b2 * x) - (b3 * y) - (b4 * z)
    c2 = (b1 * x) + (b2 * zero) + (b3 * z) - (b4 * y)
    c3 = (b1 * y) - (b2 * z) + (b3 * zero) + (b4 * x)
    c4 = (b1 * z) + (b2 * y) - (b3 * x) + (b4 * zero)
    n2, n3, n4 = -b2, -b3, -b4
    return np.array([(c1 * n2) + (c2 * b1) + (c3 * n4) - (c4 * n3),
                     (c1 * n3) - (c2 * n4) + (c3 * b1) + (c4 * n2),
                     (c1 * n4) + (c2 * n3) - (c3 * n2) + (c4 * b1)], dtype=_F64)


def placement_quaternions(bb, tmpl_n, tmpl_c):
    """(q1, q2, ca): quaternions (cos, sin*axis) float64 and the backbone CA float32.

    ``bb`` float32 (3,3) N,CA,C; ``tmpl_n``/``tmpl_c`` float32 template atoms 0 and 2
    (template CA at the origin)."""
    bb = np.asarray(bb, dtype=_F32)
    tmpl_n = np.asarray(tmpl_n, dtype=_F32); tmpl_c = np.asarray(tmpl_c, dtype=_F32)
    rel = (bb - bb[1]).astype(_F32)
    q1 = _quat(_unit32(_cross32(tmpl_n, rel[0])), _angle(rel[0], tmpl_n))
    n1 = _rotate(q1, tmpl_n).astype(_F32)
    c1 = _rotate(q1, tmpl_c).astype(_F32)
    plane_bb = _cross32(rel[0], rel[2])
    plane_t = _cross32(n1, c1)
    q2 = _quat(_unit32(_cross32(plane_t, plane_bb)), _angle(plane_t, plane_bb))
    return q1, q2, bb[1].copy()
