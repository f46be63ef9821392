"""ctypes binding of libmlegs_b200.so (the C ABI in include/mlegs_b200.h).

Fails loudly: there is no Python/CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libmlegs_b200.so")


class MlegsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[mlegs_b200 error {code}] {msg}")
        self.code = code
        self.msg = msg


class Params(C.Structure):
    """mlegs_params -- the globals of modules/mlegs_base.f90 that the hot path reads."""
    _fields_ = [("nr", C.c_int), ("np", C.c_int), ("nz", C.c_int),
                ("nrchop", C.c_int), ("npchop", C.c_int), ("nzchop", C.c_int),
                ("ell", C.c_double), ("zlen", C.c_double), ("visc", C.c_double),
                ("hyperpow", C.c_int), ("hypervisc", C.c_double), ("is_svv", C.c_int),
                ("svv_cutoff", C.c_double), ("svv_target", C.c_double),
                ("svv_strength", C.c_double), ("svv_relax", C.c_double)]


class Field(C.Structure):
    """mlegs_field -- mirror of type(scalar), modules/mlegs_scalar.f90:15-49."""
    _fields_ = [("e", C.c_void_p),
                ("glb_sz", C.c_int * 3), ("loc_sz", C.c_int * 3), ("loc_st", C.c_int * 3),
                ("axis_comm", C.c_int * 3),
                ("ln", C.c_double),
                ("nrchop_offset", C.c_int), ("npchop_offset", C.c_int), ("nzchop_offset", C.c_int),
                ("space", C.c_char * 4)]


_lib = None

_P = C.POINTER
_SIGS = {
    "mlegs_b200_last_error": (C.c_char_p, []),
    "mlegs_b200_version": (C.c_int, []),
    "mlegs_b200_set_stream": (C.c_int, [C.c_void_p]),
    "mlegs_b200_device_sync": (C.c_int, []),
    "mlegs_b200_launch_count": (C.c_longlong, [C.c_int]),
    "mlegs_b200_prof_enable": (C.c_int, [C.c_int]),
    "mlegs_b200_prof_report": (C.c_int, [C.c_char_p, C.c_size_t]),
    "mlegs_b200_dmma_peak": (C.c_int, [_P(C.c_double)]),
    "mlegs_b200_tfm_tables": (C.c_int, [_P(Params)] + [C.c_void_p] * 9),
    "mlegs_b200_tfm_tables_cached": (C.c_int, [_P(Params), C.c_char_p] + [C.c_void_p] * 9 + [_P(C.c_int)]),
    "mlegs_b200_init": (C.c_int, [_P(Params)] + [C.c_void_p] * 6 + [C.c_int, C.c_int]),
    "mlegs_b200_finalize": (C.c_int, []),
    "mlegs_b200_update_params": (C.c_int, [_P(Params)]),
    "mlegs_b200_field_alloc": (C.c_int, [_P(Field), C.c_char_p]),
    "mlegs_b200_use_managed": (C.c_int, [C.c_int]),
    "mlegs_b200_field_free": (C.c_int, [_P(Field)]),
    "mlegs_b200_field_copy": (C.c_int, [_P(Field), _P(Field)]),
    "mlegs_b200_field_zero": (C.c_int, [_P(Field)]),
    "mlegs_b200_field_upload": (C.c_int, [_P(Field), C.c_void_p]),
    "mlegs_b200_field_download": (C.c_int, [_P(Field), C.c_void_p]),
    "mlegs_b200_field_chop_offset": (C.c_int, [_P(Field), C.c_int, C.c_int, C.c_int]),
    "mlegs_b200_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mlegs_b200_host_unregister": (C.c_int, [C.c_void_p]),
    "mlegs_b200_trans": (C.c_int, [_P(Field), C.c_char_p]),
    "mlegs_b200_trans_many": (C.c_int, [C.c_int, C.c_void_p, C.c_char_p]),
    "mlegs_b200_trans_host": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double]),
    "mlegs_b200_trans_host_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p]),
    "mlegs_b200_exchange": (C.c_int, [_P(Field), C.c_int, C.c_int]),
    "mlegs_b200_chop": (C.c_int, [_P(Field)]),
    "mlegs_b200_dealias": (C.c_int, [_P(Field)]),
    "mlegs_b200_svv_filter": (C.c_int, [_P(Field), _P(C.c_double)]),
    "mlegs_b200_calcat0": (C.c_int, [_P(Field), C.c_void_p]),
    "mlegs_b200_calcat1": (C.c_int, [_P(Field), C.c_void_p]),
    "mlegs_b200_zeroat1": (C.c_int, [_P(Field)]),
    "mlegs_b200_fftreat": (C.c_int, [_P(Field)]),
    "mlegs_b200_delsqp": (C.c_int, [_P(Field)]),
    "mlegs_b200_idelsqp": (C.c_int, [_P(Field)]),
    "mlegs_b200_xxdx": (C.c_int, [_P(Field)]),
    "mlegs_b200_del2h": (C.c_int, [_P(Field)]),
    "mlegs_b200_del2": (C.c_int, [_P(Field)]),
    "mlegs_b200_idel2": (C.c_int, [_P(Field), C.c_int, C.c_double]),
    "mlegs_b200_ihelm": (C.c_int, [_P(Field), C.c_double]),
    "mlegs_b200_helmp": (C.c_int, [_P(Field), C.c_int, C.c_double, C.c_double]),
    "mlegs_b200_ihelmp": (C.c_int, [_P(Field), C.c_int, C.c_double, C.c_double]),
    "mlegs_b200_solve_cache": (C.c_int, [C.c_int]),
    "mlegs_b200_fefe": (C.c_int, [_P(Field), _P(Field), C.c_double]),
    "mlegs_b200_febe": (C.c_int, [_P(Field), _P(Field), C.c_double]),
    "mlegs_b200_abcn": (C.c_int, [_P(Field)] * 4 + [C.c_double]),
    "mlegs_b200_abab": (C.c_int, [_P(Field)] * 4 + [C.c_double, C.c_int]),
    "mlegs_b200_helm": (C.c_int, [_P(Field), C.c_double]),
    "mlegs_b200_vecprod": (C.c_int, [_P(Field)] * 6),
    "mlegs_b200_vec2tp": (C.c_int, [_P(Field)] * 5),
    "mlegs_b200_tp2vec": (C.c_int, [_P(Field)] * 5),
    "mlegs_b200_tp2curlvec": (C.c_int, [_P(Field)] * 5),
    "mlegs_b200_axpby": (C.c_int, [_P(Field), C.c_double, _P(Field), C.c_double]),
    "mlegs_b200_is_finite": (C.c_int, [_P(Field), _P(C.c_int)]),
    "mlegs_b200_gauss_vortices": (C.c_int, [_P(Field), C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                            C.c_double, C.c_ulonglong]),
    "mlegs_b200_fill_physical": (C.c_int, [_P(Field), C.c_double, C.c_double]),
    "mlegs_b200_qvort_dist_tp": (C.c_int, [_P(Field), _P(Field), C.c_double, C.c_double, C.c_ulonglong]),
    "mlegs_b200_vort_mag": (C.c_int, [_P(Field)] * 6),
    "mlegs_b200_msave": (C.c_int, [_P(Field), C.c_char_p, C.c_int, C.c_int]),
    "mlegs_b200_mload": (C.c_int, [C.c_char_p, _P(Field), C.c_int, C.c_int]),
    "mlegs_b200_msave_part": (C.c_int, [_P(Field), C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mlegs_b200_mload_part": (C.c_int, [C.c_char_p, _P(Field), C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "mlegs_b200_dist_window": (C.c_int, [_P(C.c_void_p), _P(C.c_size_t), C.c_void_p]),
    "mlegs_b200_dist_attach": (C.c_int, [C.c_void_p]),
    "mlegs_b200_dist_detach": (C.c_int, []),
    "mlegs_b200_dist_allreduce": (C.c_int, [C.c_void_p, C.c_int]),
    "mlegs_b200_dist_put_map": (C.c_int, [C.c_int] * 6 + [C.c_void_p, C.c_void_p]),
    "mlegs_b200_dist_m_stride": (C.c_int, [_P(Field), _P(C.c_int)]),
    "mlegs_b200_dist_stage_map": (C.c_int, [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p]),
}


def exported_symbols():
    return sorted(_SIGS)


def lib() -> C.CDLL:
    """Load the shared library (built in-tree by mlegs_b200/build.py); no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -m mlegs_b200.build` (or __graft_entry__.build()); "
                "mlegs_b200 has no CPU fallback")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise MlegsError(rc, lib().mlegs_b200_last_error().decode(errors="replace"))
