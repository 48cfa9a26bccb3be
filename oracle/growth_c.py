"""ORACLE - test infrastructure only.  ctypes front end of oracle/growth_c.c (the C restatement of
the growth hot path; one conformer per OpenMP thread).  Used by tests (checked against
oracle/ref_numpy.py) and by bench.py as the CPU baseline / ``--impl reference`` arm."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from mcsce_b200.chem.tables import load_tables
from mcsce_b200.system import placement_backbone

from oracle import ref_numpy as _O


def _placement_quaternions(bb, tmpl):
    """The two rotations of place_sidechain_template (libcalc.py:449-480) as quaternions, derived with the
    NumPy oracle's float32-BLAS helpers."""
    bb = np.asarray(bb, np.float32); tmpl = np.asarray(tmpl, np.float32)
    rel = (bb - bb[1]).astype(np.float32)
    a1 = _O.angle_between(rel[0], tmpl[0])
    rv = _O._cross32(tmpl[0], rel[0]); rvu = rv / _O._norm32(rv)
    quat = lambda ax, ang: np.array([np.cos(np.float64(ang) / 2), *(np.sin(np.float64(ang) / 2) * ax.astype(np.float64))])  # noqa: E731
    q1 = quat(rvu, a1)
    rot1 = _O.rotate_q(tmpl[[0, 2]], rvu, a1).astype(np.float32)
    cc = _O._cross32(rel[0], rel[2]); cs = _O._cross32(rot1[0], rot1[1])
    a2 = _O.angle_between(cs, cc)
    rv = _O._cross32(cs, cc); rvu = rv / _O._norm32(rv)
    return q1, quat(rvu, a2), bb[1].copy()


HERE = Path(__file__).resolve().parent
LIB = HERE / "libgrowth_c.so"
MAX_TMPL, MAX_SC, MAX_ROT, MAX_CHI, MAX_SP_ENV = 32, 28, 5, 6, 16

_ip, _up, _fp, _dp = (C.POINTER(t) for t in (C.c_int32, C.c_uint32, C.c_float, C.c_double))


class _System(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int), ("n_input", C.c_int), ("n_res", C.c_int), ("n_cls", C.c_int), ("n_types", C.c_int),
        ("cls_of_atom", _ip), ("charge", _dp), ("cls_s2", _dp), ("cls_e4", _dp), ("cls_thr", _dp),
        ("tmpl_n_atoms", _ip), ("tmpl_xyz", _fp), ("tmpl_sc_pos", _ip), ("tmpl_dih", _ip), ("tmpl_axis", _ip),
        ("tmpl_move", _up),
        ("step_kind", _ip), ("step_type", _ip), ("step_n_sc", _ip), ("step_sc_start", _ip), ("step_n_sp_env", _ip),
        ("step_sp_env", _ip), ("step_sp_off", _ip), ("step_n_sp", _ip),
        ("sp_k", _ip), ("sp_partner", _ip), ("sp_coef", _dp),
        ("n_bb_pairs", C.c_int), ("bb_i", _ip), ("bb_j", _ip), ("bb_coef", _dp), ("resnum", _ip),
        ("step_n_trials", _ip), ("step_trial_off", _ip), ("step_n_chi_lib", _ip), ("step_n_chi", _ip),
        ("chi_mean", _dp), ("chi_sigma", _dp), ("prob", _dp), ("place_q1", _dp), ("place_q2", _dp), ("place_ca", _fp), ("input_xyz", _fp),
        ("n_trials_total", C.c_int),
    ]


def load():
    if not LIB.exists():
        from mcsce_b200.build import build_oracle_c
        build_oracle_c()
    lib = C.CDLL(str(LIB))
    lib.oracle_grow.restype = C.c_int
    lib.oracle_grow.argtypes = [C.POINTER(_System), C.c_int, C.c_double, C.c_double, C.c_uint64, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p]
    lib.oracle_input_energy.restype = C.c_double
    lib.oracle_input_energy.argtypes = [C.POINTER(_System)]
    lib.oracle_max_threads.restype = C.c_int
    lib.oracle_rescore.restype = C.c_double
    lib.oracle_rescore.argtypes = [C.POINTER(_System), C.c_void_p, C.POINTER(C.c_double)]
    lib.oracle_rescore_levels.restype = C.c_double
    lib.oracle_rescore_levels.argtypes = [C.POINTER(_System), C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
    return lib


class COracle:
    def __init__(self, sysm, backbone_coords, tables=None):
        t = tables or load_tables()
        self.lib = load()
        self.sysm = sysm
        types = sorted({st.resname for st in sysm.steps if st.kind != "skip"})
        tidx = {r: i for i, r in enumerate(types)}
        nt = max(1, len(types))
        a = {}
        a["tmpl_n_atoms"] = np.zeros(nt, np.int32); a["tmpl_xyz"] = np.zeros((nt, MAX_TMPL, 3), np.float32)
        a["tmpl_sc_pos"] = np.zeros((nt, MAX_SC), np.int32); a["tmpl_dih"] = np.zeros((nt, MAX_ROT, 4), np.int32)
        a["tmpl_axis"] = np.zeros((nt, MAX_ROT, 2), np.int32); a["tmpl_move"] = np.zeros((nt, MAX_ROT), np.uint32)
        for r, i in tidx.items():
            tm = t.templates[r]
            a["tmpl_n_atoms"][i] = tm.n_atoms; a["tmpl_xyz"][i, :tm.n_atoms] = tm.coords
            a["tmpl_sc_pos"][i, :tm.n_sc] = tm.sc_idx
            for k, chi in enumerate(tm.chis[:MAX_ROT]):
                a["tmpl_dih"][i, k] = chi["dihedral"]; a["tmpl_axis"][i, k] = chi["axis"]
                a["tmpl_move"][i, k] = sum(1 << m for m in chi["moving"])
        L = sysm.n_res
        a["step_kind"] = np.array([{"skip": 0, "rigid": 1, "grow": 2}[st.kind] for st in sysm.steps], np.int32)
        a["step_type"] = np.array([tidx.get(st.resname, 0) for st in sysm.steps], np.int32)
        a["step_n_sc"] = np.array([st.n_sc for st in sysm.steps], np.int32)
        a["step_sc_start"] = np.array([st.sc_start for st in sysm.steps], np.int32)
        a["step_n_sp_env"] = np.zeros(L, np.int32); a["step_sp_env"] = np.zeros((L, MAX_SP_ENV), np.int32)
        a["step_sp_off"] = np.zeros(L, np.int32); a["step_n_sp"] = np.zeros(L, np.int32)
        a["step_n_trials"] = np.zeros(L, np.int32); a["step_trial_off"] = np.zeros(L, np.int32)
        a["step_n_chi_lib"] = np.zeros(L, np.int32); a["step_n_chi"] = np.zeros(L, np.int32)
        ks, pas, cfs, means, sigs, probs = [], [], [], [], [], []
        off = toff = 0
        for st in sysm.steps:
            if st.kind != "grow":
                continue
            i = st.index
            a["step_n_sp_env"][i] = len(st.special_env); a["step_sp_env"][i, :len(st.special_env)] = st.special_env
            a["step_sp_off"][i] = off; a["step_n_sp"][i] = len(st.sp_k); off += len(st.sp_k)
            ks.append(st.sp_k); pas.append(st.sp_partner); cfs.append(st.sp_coef)
            T = len(st.prob)
            m = np.zeros((T, MAX_CHI)); s = np.zeros((T, MAX_CHI))
            m[:, :st.n_chi_lib] = st.chi_mean; s[:, :st.n_chi_lib] = st.chi_sigma
            means.append(m); sigs.append(s); probs.append(st.prob)
            a["step_n_trials"][i], a["step_trial_off"][i], a["step_n_chi_lib"][i], a["step_n_chi"][i] = T, toff, st.n_chi_lib, st.n_chi
            toff += T
        cat = lambda xs, dt, shape: (np.concatenate(xs).astype(dt) if xs else np.zeros(shape, dt))  # noqa: E731
        a["sp_k"] = cat(ks, np.int32, (0,)); a["sp_partner"] = cat(pas, np.int32, (0,)); a["sp_coef"] = cat(cfs, np.float64, (0, 4))
        a["chi_mean"] = cat(means, np.float64, (0, MAX_CHI)); a["chi_sigma"] = cat(sigs, np.float64, (0, MAX_CHI))
        a["prob"] = cat(probs, np.float64, (0,))
        a["cls_of_atom"] = sysm.cls_of_atom.astype(np.int32)
        a["charge"] = (sysm.charge if "coulomb" in sysm.terms else np.zeros_like(sysm.charge)).astype(np.float64)
        a["cls_s2"], a["cls_e4"], a["cls_thr"] = (np.ascontiguousarray(x, np.float64) for x in (sysm.cls_s2, sysm.cls_e4, sysm.cls_thr))
        bi, bj, bc = sysm.bb_pairs
        a["bb_i"], a["bb_j"], a["bb_coef"] = bi.astype(np.int32), bj.astype(np.int32), np.ascontiguousarray(bc, np.float64)
        a["resnum"] = sysm.resnums.astype(np.int32)
        bb = np.asarray(backbone_coords, np.float32)
        ncac = placement_backbone(bb, sysm)
        q1 = np.zeros((L, 4)); q2 = np.zeros((L, 4)); ca = np.zeros((L, 3), np.float32)
        for st in sysm.steps:
            if st.kind != "skip":
                tm = t.templates[st.resname]
                q1[st.index], q2[st.index], ca[st.index] = _placement_quaternions(ncac[st.index], tm.coords)
        a["place_q1"], a["place_q2"], a["place_ca"] = q1, q2, ca
        a["input_xyz"] = np.ascontiguousarray(bb, np.float32)
        a = {k: np.ascontiguousarray(v) for k, v in a.items()}
        d = _System()
        d.n_atoms, d.n_input, d.n_res, d.n_cls, d.n_types = sysm.n_atoms, sysm.n_input, L, sysm.n_cls, len(types)
        d.n_bb_pairs, d.n_trials_total = len(bi), toff
        ftype = dict(_System._fields_)
        for k, v in a.items():
            setattr(d, k, v.ctypes.data_as(ftype[k]))
        self._arrays, self.desc = a, d
        self.n_trials_total = toff
        self.trial_off = a["step_trial_off"]

    def input_energy(self):
        return float(self.lib.oracle_input_energy(C.byref(self.desc)))

    def rescore(self, coords):
        """(total energy, min distance/clash-distance ratio) of one finished conformer, growth-order (N,3) float32."""
        xyz = np.ascontiguousarray(coords, np.float32)
        assert xyz.shape == (self.sysm.n_atoms, 3)
        r = C.c_double(0.0)
        e = self.lib.oracle_rescore(C.byref(self.desc), xyz.ctypes.data_as(C.c_void_p), C.byref(r))
        return float(e), float(r.value)

    def rescore_levels(self, coords):
        """(total, ratio, per-residue energies): like ``rescore`` plus the energy of every grown residue's side chain
        against everything placed before it - what the growth scored for the trial it selected."""
        xyz = np.ascontiguousarray(coords, np.float32)
        assert xyz.shape == (self.sysm.n_atoms, 3)
        r = C.c_double(0.0)
        lev = np.zeros(self.sysm.n_res)
        e = self.lib.oracle_rescore_levels(C.byref(self.desc), xyz.ctypes.data_as(C.c_void_p), C.byref(r),
                                           lev.ctypes.data_as(C.c_void_p))
        return float(e), float(r.value), lev

    def grow(self, n_conf, beta, seed=0, noise_deg=0.5, chis=None, uniforms=None, dump=False, n_threads=0, oris=None):
        """``oris`` (n_conf, n_trials_total, MAX_ROT) rad: dihedrals to use instead of the float32 re-measurement (the
        reference's recorded values; NaN = measure).  ``dump`` also returns the values used as ``ori``."""
        N = self.sysm.n_atoms
        coords = np.zeros((n_conf, N, 3), np.float32); energy = np.zeros(n_conf); ok = np.zeros(n_conf, np.uint8)
        vp = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)  # noqa: E731
        ch = None if chis is None else np.ascontiguousarray(chis, np.float64)
        un = None if uniforms is None else np.ascontiguousarray(uniforms, np.float64)
        te = sel = txyz = ori_out = None
        ori_in = None if oris is None else np.ascontiguousarray(oris, np.float64)
        if dump:
            ori_out = np.full((n_conf, self.n_trials_total, MAX_ROT), np.nan)
            te = np.full((n_conf, self.n_trials_total), np.nan); sel = np.full((n_conf, self.sysm.n_res), -2, np.int32)
            txyz = np.zeros((self.n_trials_total, MAX_SC, 3), np.float32)
        self.lib.oracle_grow(C.byref(self.desc), n_conf, float(beta), float(noise_deg), int(seed), vp(ch), vp(un),
                             vp(coords), vp(energy), vp(ok), vp(te), vp(sel), vp(txyz), int(n_threads), vp(ori_in), vp(ori_out))
        out = dict(coords=coords, energy=energy, ok=ok)
        if dump:
            out.update(trial_energy=te, selected=sel, trial_xyz=txyz, ori=ori_out)
        return out
