"""ctypes binding of ``libmcsce_b200.so`` (the C ABI in ``include/mcsce_b200.h``).

There is no CPU fallback: importing the engine without the built CUDA library, or on a
machine without a CUDA device, raises.  Build with ``python -m mcsce_b200.build`` (nvcc,
sm_100a) - the library is kept in-tree next to this file.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

LIB_PATH = Path(os.environ.get("MCSCE_B200_LIB") or (Path(__file__).resolve().parent / "libmcsce_b200.so"))
if os.environ.get("MCSCE_B200_LIB"):          # build variants for timing experiments (tools/)
    LIB_PATH = Path(os.environ["MCSCE_B200_LIB"])

MAX_CLS, MAX_SP_ENV, MAX_CHI, MAX_CHI_ROT, MAX_TMPL, MAX_SC = 64, 16, 6, 5, 32, 28

_i32p, _u32p, _f32p, _f64p = (C.POINTER(t) for t in (C.c_int32, C.c_uint32, C.c_float, C.c_double))


class SystemDesc(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int32), ("n_input", C.c_int32), ("n_res", C.c_int32), ("n_cls", C.c_int32),
        ("n_types", C.c_int32), ("n_sp_total", C.c_int32),
        ("slot_of_atom", _i32p), ("cls_of_atom", _i32p), ("charge", _f64p), ("resnum", _i32p),
        ("cls_start", _i32p), ("cls_active", _i32p), ("cls_s2", _f64p), ("cls_e4", _f64p), ("cls_thr", _f64p),
        ("tmpl_n_atoms", _i32p), ("tmpl_xyz", _f32p), ("tmpl_sc_pos", _i32p), ("tmpl_n_chi", _i32p),
        ("tmpl_axis", _i32p), ("tmpl_move", _u32p), ("tmpl_ori", _f64p),
        ("step_kind", _i32p), ("step_type", _i32p), ("step_n_sc", _i32p), ("step_sc_start", _i32p),
        ("step_n_sp_env", _i32p), ("step_sp_env", _i32p), ("step_sp_off", _i32p), ("step_n_sp", _i32p),
        ("sp_k", _i32p), ("sp_partner", _i32p), ("sp_coef", _f64p),
        ("n_bb_pairs", C.c_int32), ("bb_i", _i32p), ("bb_j", _i32p), ("bb_coef", _f64p),
    ]


class BackboneDesc(C.Structure):
    _fields_ = [
        ("n_trials_total", C.c_int32),
        ("step_n_trials", _i32p), ("step_trial_off", _i32p), ("step_n_chi_lib", _i32p), ("step_n_chi", _i32p),
        ("chi_mean", _f64p), ("chi_sigma", _f64p), ("prob", _f64p),
        ("place_q1", _f64p), ("place_q2", _f64p), ("place_ca", _f32p), ("input_xyz", _f32p),
    ]


class GrowOpts(C.Structure):
    _fields_ = [
        ("n_conf", C.c_int32), ("conf_id0", C.c_int64), ("seed", C.c_uint64), ("beta", C.c_double),
        ("noise_deg", C.c_double), ("d_chis", C.c_void_p), ("d_uniform", C.c_void_p),
        ("d_trial_energy", C.c_void_p), ("d_chis_out", C.c_void_p), ("d_selected", C.c_void_p),
        ("env_split", C.c_int32), ("no_dedup", C.c_int32), ("no_graph", C.c_int32), ("no_screen", C.c_int32),
        ("d_ori", C.c_void_p), ("hpmf", C.c_void_p),
    ]


SASA_MAX_BONDS = 4


class HpmfDesc(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int32), ("n_typed", C.c_int32), ("n_types", C.c_int32),
        ("typed_atom", _i32p), ("type_of", _i32p), ("bonded", _i32p),
        ("type_R", _f64p), ("type_P", _f64p), ("type_is_carbon", C.POINTER(C.c_uint8)),
        ("r_solvent", C.c_double), ("pij_bonded", C.c_double), ("pij_non_bonded", C.c_double),
        ("hpmf_c", C.c_double * 3), ("hpmf_w", C.c_double * 3), ("hpmf_h", C.c_double * 3),
        ("sa_threshold", C.c_double),
    ]


#: every symbol include/mcsce_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mcsce_last_error": (C.c_char_p, []),
    "mcsce_version": (C.c_int, []),
    "mcsce_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "mcsce_destroy": (C.c_int, [C.c_void_p]),
    "mcsce_set_system": (C.c_int, [C.c_void_p, C.POINTER(SystemDesc)]),
    "mcsce_set_backbone": (C.c_int, [C.c_void_p, C.POINTER(BackboneDesc)]),
    "mcsce_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "mcsce_grow": (C.c_int, [C.c_void_p, C.POINTER(GrowOpts), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_void_p]),
    "mcsce_grow_host": (C.c_int, [C.c_void_p, C.POINTER(GrowOpts), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "mcsce_place_trials": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcsce_peer_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "mcsce_peer_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    "mcsce_peer_close": (C.c_int, [C.c_int, C.c_void_p]),
    "mcsce_peer_free": (C.c_int, [C.c_int, C.c_void_p]),
    "mcsce_place_trials_replay": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcsce_score_trials": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "mcsce_score_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "mcsce_pack_env": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcsce_select": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcsce_pick": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcsce_input_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcsce_init_env": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mcsce_rotate_template": (C.c_int, [C.c_int, C.c_int, _f32p, C.c_int, _i32p, _u32p, _f64p, C.c_int, _f64p, _f32p]),
    "mcsce_place_template": (C.c_int, [C.c_int, C.c_int, _f32p, _f64p, _f64p, _f32p, _f64p]),
    "mcsce_write_pdbs": (C.c_int, [C.c_int, C.c_int, _f32p, C.c_void_p, _i32p, C.c_char_p, C.c_int, C.c_int, C.c_char_p,
                                   C.c_int, _i32p]),
    "mcsce_profile": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mcsce_hpmf_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(HpmfDesc)]),
    "mcsce_hpmf_destroy": (C.c_int, [C.c_void_p]),
    "mcsce_sasa": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcsce_hpmf_energy": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcsce_hpmf_trials": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "mcsce_hpmf_carry": (C.c_int, [C.c_void_p, C.c_int]),
    "mcsce_hpmf_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mcsce_hpmf_last_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "mcsce_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]),
}

_lib = None


def load_library():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            "%s is missing: the CUDA extension has not been built (python -m mcsce_b200.build). "
            "mcsce_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class McsceError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise McsceError(load_library().mcsce_last_error().decode(errors="replace"))
