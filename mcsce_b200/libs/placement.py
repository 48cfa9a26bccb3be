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


# ---------------------------------------------------------------------------------------------------------------
# All residues of a backbone at once.  Same arithmetic, value by value: the element-wise float32 / float64 steps are
# NumPy array operations (one rounding per operation, like the scalar code above), and the two reductions whose
# rounding belongs to the BLAS (snrm2, sdot) are still BLAS calls, one per vector.
# ---------------------------------------------------------------------------------------------------------------
def _norms32(v):
    return np.array([_blas.snrm2(row) for row in v], dtype=_F32)


def _unit32_rows(v):
    v = np.ascontiguousarray(v, dtype=_F32)
    return (v / _norms32(v)[:, None]).astype(_F32)


def _angle_rows(v1, v2):
    u1, u2 = _unit32_rows(v1), _unit32_rows(v2)
    d = np.array([_blas.sdot(a, b) for a, b in zip(u1, u2)], dtype=_F32).astype(_F64)
    return np.arccos(np.minimum(np.maximum(d, _F64(-1.0)), _F64(1.0)))


def _cross32_rows(a, b):
    out = np.empty(a.shape, dtype=_F32)
    out[:, 0] = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    out[:, 1] = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
    out[:, 2] = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    return out


def _quat_rows(axis_unit32, angle):
    half = angle.astype(_F64) / 2
    s = np.sin(half)
    a = axis_unit32.astype(_F64)
    return np.stack([np.cos(half), s * a[:, 0], s * a[:, 1], s * a[:, 2]], axis=1)


def _rotate_rows(q, p):
    b1, b2, b3, b4 = (q[:, k] for k in range(4))
    x, y, z = (p[:, k].astype(_F64) for k in range(3))
    zero = _F64(0.0)
    c1 = (b1 * zero) - (b2 * x) - (b3 * y) - (b4 * z)
    c2 = (b1 * x) + (b2 * zero) + (b3 * z) - (b4 * y)
    c3 = (b1 * y) - (b2 * z) + (b3 * zero) + (b4 * x)
    c4 = (b1 * z) + (b2 * y) - (b3 * x) + (b4 * zero)
    n2, n3, n4 = -b2, -b3, -b4
    return np.stack([(c1 * n2) + (c2 * b1) + (c3 * n4) - (c4 * n3),
                     (c1 * n3) - (c2 * n4) + (c3 * b1) + (c4 * n2),
                     (c1 * n4) + (c2 * n3) - (c3 * n2) + (c4 * b1)], axis=1)


def placement_quaternions_batch(bb, tmpl_n, tmpl_c):
    """``placement_quaternions`` for L residues: ``bb`` (L,3,3) float32, ``tmpl_n``/``tmpl_c`` (L,3) float32 ->
    ``q1`` (L,4), ``q2`` (L,4) float64, ``ca`` (L,3) float32; bit-identical to L scalar calls."""
    bb = np.ascontiguousarray(bb, dtype=_F32)
    tmpl_n = np.ascontiguousarray(tmpl_n, dtype=_F32); tmpl_c = np.ascontiguousarray(tmpl_c, dtype=_F32)
    rel = (bb - bb[:, 1:2]).astype(_F32)
    rel0, rel2 = np.ascontiguousarray(rel[:, 0]), np.ascontiguousarray(rel[:, 2])
    q1 = _quat_rows(_unit32_rows(_cross32_rows(tmpl_n, rel0)), _angle_rows(rel0, tmpl_n))
    n1 = _rotate_rows(q1, tmpl_n).astype(_F32)
    c1 = _rotate_rows(q1, tmpl_c).astype(_F32)
    plane_bb = _cross32_rows(rel0, rel2)
    plane_t = _cross32_rows(n1, c1)
    q2 = _quat_rows(_unit32_rows(_cross32_rows(plane_t, plane_bb)), _angle_rows(plane_t, plane_bb))
    return q1, q2, bb[:, 1].copy()
