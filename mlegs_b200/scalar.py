"""Host mirror of type(scalar) and of the operator interfaces of modules/mlegs_scalar.f90.

Every function forwards to the C-ABI entry point of the same name; data stay in HBM.
Names, argument meaning and error text follow the reference's interfaces
(modules/mlegs_scalar.f90:115-428) so tests read like the reference's tutorials.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Field, check


class Scalar:
    """type(scalar): a distributed complex field resident on the device."""

    def __init__(self, space: str = "PPP"):
        self.f = Field()
        check(_lib.lib().mlegs_b200_field_alloc(C.byref(self.f), space.encode()))

    # -- metadata ------------------------------------------------------------------------
    @property
    def space(self) -> str:
        return self.f.space.decode()

    @space.setter
    def space(self, v: str):
        self.f.space = v.encode()

    @property
    def ln(self) -> float:
        return self.f.ln

    @ln.setter
    def ln(self, v: float):
        self.f.ln = float(v)

    @property
    def loc_sz(self):
        return tuple(self.f.loc_sz)

    @property
    def loc_st(self):
        return tuple(self.f.loc_st)

    @property
    def glb_sz(self):
        return tuple(self.f.glb_sz)

    def chop_offset(self, iof1: int, iof2: int = 0, iof3: int = 0):
        check(_lib.lib().mlegs_b200_field_chop_offset(C.byref(self.f), iof1, iof2, iof3))

    # -- data movement ------------------------------------------------------------------------
    def upload(self, host: np.ndarray) -> "Scalar":
        a = np.asfortranarray(host, dtype=np.complex128)
        if a.shape != self.loc_sz:
            raise ValueError(f"scalar upload: shape {a.shape} != local size {self.loc_sz}")
        check(_lib.lib().mlegs_b200_field_upload(C.byref(self.f), a.ctypes.data_as(C.c_void_p)))
        return self

    @property
    def m_stride(self) -> int:
        """Local azimuthal column j holds global m = loc_st[1] + m_stride * j (the number of ranks when m is distributed:
        cyclic ownership, include/mlegs_b200.h; 1 otherwise)."""
        v = C.c_int(1)
        check(_lib.lib().mlegs_b200_dist_m_stride(C.byref(self.f), C.byref(v)))
        return v.value

    def global_slices(self):
        """Index of this rank's block inside the global array."""
        st, sz, ms = self.loc_st, self.loc_sz, self.m_stride
        return (slice(st[0], st[0] + sz[0]), slice(st[1], st[1] + sz[1] * ms, ms), slice(st[2], st[2] + sz[2]))

    def upload_global(self, glb: np.ndarray) -> "Scalar":
        """disassemble (dist:205-368) without the gather: every rank holds the global array and keeps its slab."""
        return self.upload(glb[self.global_slices()])

    def download(self) -> np.ndarray:
        out = np.empty(self.loc_sz, dtype=np.complex128, order="F")
        check(_lib.lib().mlegs_b200_field_download(C.byref(self.f), out.ctypes.data_as(C.c_void_p)))
        return out

    def copy(self) -> "Scalar":
        o = Scalar.__new__(Scalar)
        o.f = Field()
        check(_lib.lib().mlegs_b200_field_copy(C.byref(o.f), C.byref(self.f)))
        return o

    def assign(self, other: "Scalar"):
        """this = that (scalar_copy, submodules/mlegs_scalar_init.f90:106-140)."""
        check(_lib.lib().mlegs_b200_field_copy(C.byref(self.f), C.byref(other.f)))

    def zero(self):
        check(_lib.lib().mlegs_b200_field_zero(C.byref(self.f)))

    def exchange(self, axis_old: int, axis_new: int):
        check(_lib.lib().mlegs_b200_exchange(C.byref(self.f), axis_old, axis_new))

    def dealloc(self):
        if self.f.e:
            check(_lib.lib().mlegs_b200_field_free(C.byref(self.f)))

    def __del__(self):
        try:
            self.dealloc()
        except Exception:
            pass


def _l():
    return _lib.lib()


def trans(s: Scalar, space: str):
    check(_l().mlegs_b200_trans(C.byref(s.f), space.encode()))


def trans_many(scalars, space: str):
    """trans() of several scalars in the same state (e.g. the components of a vector field): one launch per stage."""
    n = len(scalars)
    ptrs = (C.POINTER(Field) * n)(*[C.pointer(s.f) for s in scalars])
    check(_l().mlegs_b200_trans_many(n, ptrs, space.encode()))


def trans_host(host_e: np.ndarray, from_space: str, to_space: str, ln: float = 0.0):
    """Reference-facing call on a host array (the Fortran s%e): H2D + trans + D2H."""
    assert host_e.flags.f_contiguous and host_e.dtype == np.complex128
    check(_l().mlegs_b200_trans_host(host_e.ctypes.data_as(C.c_void_p), from_space.encode(), to_space.encode(), ln))


def trans_host_batch(host_arrays, from_space: str, to_space: str, ln=None):
    """trans_host for several independent host arrays, pipelined over PCIe (H2D / transform / D2H overlap)."""
    n = len(host_arrays)
    for a in host_arrays:
        assert a.flags.f_contiguous and a.dtype == np.complex128
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in host_arrays])
    lnv = None
    if ln is not None:
        lnv = (C.c_double * n)(*[float(v) for v in ln])
    check(_l().mlegs_b200_trans_host_batch(n, ptrs, from_space.encode(), to_space.encode(), lnv))


def chop(s): check(_l().mlegs_b200_chop(C.byref(s.f)))
def dealias(s): check(_l().mlegs_b200_dealias(C.byref(s.f)))


def svv_filter(s, gain: float) -> float:
    g = C.c_double(gain)
    check(_l().mlegs_b200_svv_filter(C.byref(s.f), C.byref(g)))
    return g.value


def _calcat(fn, s):
    out = np.zeros(s.glb_sz[2], dtype=np.complex128)
    check(fn(C.byref(s.f), out.ctypes.data_as(C.c_void_p)))
    return out


def calcat0(s): return _calcat(_l().mlegs_b200_calcat0, s)
def calcat1(s): return _calcat(_l().mlegs_b200_calcat1, s)
def zeroat1(s): check(_l().mlegs_b200_zeroat1(C.byref(s.f)))
def fftreat(s): check(_l().mlegs_b200_fftreat(C.byref(s.f)))
def delsqp(s): check(_l().mlegs_b200_delsqp(C.byref(s.f)))
def idelsqp(s): check(_l().mlegs_b200_idelsqp(C.byref(s.f)))
def xxdx(s): check(_l().mlegs_b200_xxdx(C.byref(s.f)))
def del2h(s): check(_l().mlegs_b200_del2h(C.byref(s.f)))
def del2(s): check(_l().mlegs_b200_del2(C.byref(s.f)))


def idel2(s, preln=None):
    if preln is None:
        check(_l().mlegs_b200_idel2(C.byref(s.f), 0, 0.0))
    else:
        check(_l().mlegs_b200_idel2(C.byref(s.f), 1, float(preln)))


def ihelm(s, alpha): check(_l().mlegs_b200_ihelm(C.byref(s.f), alpha))
def helmp(s, power, alpha, beta): check(_l().mlegs_b200_helmp(C.byref(s.f), power, alpha, beta))
def ihelmp(s, power, alpha, beta): check(_l().mlegs_b200_ihelmp(C.byref(s.f), power, alpha, beta))
def solve_cache(on: bool): check(_l().mlegs_b200_solve_cache(int(on)))
def fefe(s, nl, dt): check(_l().mlegs_b200_fefe(C.byref(s.f), C.byref(nl.f), dt))
def febe(s, nl, dt): check(_l().mlegs_b200_febe(C.byref(s.f), C.byref(nl.f), dt))


def abcn(s, s_p, nl, nl_p, dt):
    check(_l().mlegs_b200_abcn(C.byref(s.f), C.byref(s_p.f), C.byref(nl.f), C.byref(nl_p.f), dt))


def abab(s, s_p, nl, nl_p, dt, is_2nd_svis_p=False):
    check(_l().mlegs_b200_abab(C.byref(s.f), C.byref(s_p.f), C.byref(nl.f), C.byref(nl_p.f), dt, int(is_2nd_svis_p)))


def helm(s, alpha): check(_l().mlegs_b200_helm(C.byref(s.f), alpha))


def vecprod(vr, vp, vz, ur, up, uz):
    check(_l().mlegs_b200_vecprod(C.byref(vr.f), C.byref(vp.f), C.byref(vz.f), C.byref(ur.f), C.byref(up.f),
                                  C.byref(uz.f)))


def vec2tp(vr, vp, vz, psi, chi):
    check(_l().mlegs_b200_vec2tp(C.byref(vr.f), C.byref(vp.f), C.byref(vz.f), C.byref(psi.f), C.byref(chi.f)))


def tp2vec(psi, chi, vr, vp, vz):
    check(_l().mlegs_b200_tp2vec(C.byref(psi.f), C.byref(chi.f), C.byref(vr.f), C.byref(vp.f), C.byref(vz.f)))


def tp2curlvec(psi, chi, wr, wp, wz):
    check(_l().mlegs_b200_tp2curlvec(C.byref(psi.f), C.byref(chi.f), C.byref(wr.f), C.byref(wp.f), C.byref(wz.f)))


def gauss_vortices(s, centres, mul=1.0, div=1.0, ran_noise=0.0, seed=0):
    """Sum of Gaussian vortices (-exp(-d^2)*mul/div/(1-x)^2) written into the PPP scalar s on the device."""
    xo = np.ascontiguousarray([c[0] for c in centres], dtype=np.float64)
    yo = np.ascontiguousarray([c[1] for c in centres], dtype=np.float64)
    check(_l().mlegs_b200_gauss_vortices(C.byref(s.f), len(centres), xo.ctypes.data_as(C.c_void_p),
                                         yo.ctypes.data_as(C.c_void_p), mul, div, ran_noise, seed))


def fill_physical(s, re, im): check(_l().mlegs_b200_fill_physical(C.byref(s.f), re, im))


def qvort_dist_tp(psi, chi, q=1.0, ran_noise=0.0, seed=0):
    check(_l().mlegs_b200_qvort_dist_tp(C.byref(psi.f), C.byref(chi.f), q, ran_noise, seed))


def vort_mag(psi, chi, wr, wp, wz, vormag):
    check(_l().mlegs_b200_vort_mag(C.byref(psi.f), C.byref(chi.f), C.byref(wr.f), C.byref(wp.f), C.byref(wz.f),
                                   C.byref(vormag.f)))


def msave(s, fn: str, is_binary: bool = False, is_global: bool = True):
    """msave_scalar (submodules/mlegs_scalar_io.f90:6-115); several ranks fill one global file cooperatively."""
    check(_l().mlegs_b200_msave(C.byref(s.f), str(fn).encode(), int(is_binary), int(is_global)))


def mload(fn: str, s, is_binary: bool = False, is_global: bool = True):
    """mload_scalar (submodules/mlegs_scalar_io.f90:117-250); every rank reads its own slab."""
    check(_l().mlegs_b200_mload(str(fn).encode(), C.byref(s.f), int(is_binary), int(is_global)))


def axpby(y, a, x, b): check(_l().mlegs_b200_axpby(C.byref(y.f), a, C.byref(x.f), b))


def is_finite(s) -> bool:
    v = C.c_int(0)
    check(_l().mlegs_b200_is_finite(C.byref(s.f), C.byref(v)))
    return bool(v.value)


def device_sync(): check(_l().mlegs_b200_device_sync())
def set_stream(ptr): check(_l().mlegs_b200_set_stream(C.c_void_p(ptr)))
def launch_count(reset=False) -> int: return int(_l().mlegs_b200_launch_count(int(reset)))


def dmma_peak() -> float:
    v = C.c_double(0.0)
    check(_l().mlegs_b200_dmma_peak(C.byref(v)))
    return v.value


def prof_enable(on: bool = True): check(_l().mlegs_b200_prof_enable(int(on)))


def prof_report() -> dict:
    import json
    buf = C.create_string_buffer(1 << 16)
    check(_l().mlegs_b200_prof_report(buf, len(buf)))
    return json.loads(buf.value.decode())
