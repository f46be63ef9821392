"""Host driver mirroring src/apps/vortical_flow_3d.f90 on top of the device operators.

The application code (initial condition, advection right-hand side, Richardson bootstrap, ABCN loop)
stays on the host exactly as in the reference program; every field operation it performs is one of the
C-ABI entry points, so fields never leave HBM inside the time loop.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import scalar as ms
from .kit import TfmKit
from .scalar import Scalar


def qvort_dist_tp(kit: TfmKit, q: float = 1.0, ran_noise: float = 0.0, seed: int = 0):
    """apps/vortical_flow_3d.f90:258-326, filled on the device slab by slab (mlegs_b200_qvort_dist_tp); the reference
    assembles a global array on the host and scatters it."""
    psi, chi = Scalar("FFF"), Scalar("FFF")
    ms.qvort_dist_tp(psi, chi, q, ran_noise, seed)
    return psi, chi


def uniform_z_fld(kit: TfmKit, b: float = -0.5) -> Scalar:
    """apps/vortical_flow_3d.f90:328-351."""
    uz = Scalar("PPP")
    ms.fill_physical(uz, b, b)
    return uz


def vort_mag(st_or_psi, chi=None) -> Scalar:
    """The vorticity-magnitude field of save_vort_mag (apps/vortical_flow_3d.f90:411-447), PPP, on the device."""
    psi = st_or_psi.psi if chi is None else st_or_psi
    chi = st_or_psi.chi if chi is None else chi
    wr, wp, wz, mag = (Scalar("PPP") for _ in range(4))
    ms.vort_mag(psi, chi, wr, wp, wz, mag)
    return mag


@dataclass
class VortexState:
    psi: Scalar
    chi: Scalar
    nlpsi: Scalar
    nlchi: Scalar
    psi_prev: Scalar
    chi_prev: Scalar
    nlpsi_prev: Scalar
    nlchi_prev: Scalar
    uz: Scalar
    work: list
    gain_psi: float = 0.0
    gain_chi: float = 0.0


def advection_rhs(psi, chi, nlpsi, nlchi, uz, work):
    """apps/vortical_flow_3d.f90:353-395; `work` holds six PPP scalars (vr,vp,vz,wr,wp,wz)."""
    vr, vp, vz, wr, wp, wz = work
    ms.dealias(psi)
    ms.dealias(chi)
    for f in work:
        f.space = "PPP"
    ms.tp2vec(psi, chi, vr, vp, vz)
    ms.axpby(vz, 1.0, uz, 1.0)                 # vz%e = vz%e + uz%e
    ms.tp2curlvec(psi, chi, wr, wp, wz)
    ms.vecprod(vr, vp, vz, wr, wp, wz)
    ms.vec2tp(vr, vp, vz, nlpsi, nlchi)
    nlpsi.ln = 0.0
    nlchi.ln = 0.0
    ms.dealias(nlpsi)
    ms.dealias(nlchi)


def bootstrap(kit: TfmKit, dt: float, psi: Scalar, chi: Scalar, uz: Scalar) -> VortexState:
    """Richardson-extrapolated FEBE first step; apps/vortical_flow_3d.f90:116-147."""
    work = [Scalar("PPP") for _ in range(6)]
    nlpsi, nlchi = Scalar("FFF"), Scalar("FFF")
    advection_rhs(psi, chi, nlpsi, nlchi, uz, work)
    psi_prev, nlpsi_prev = psi.copy(), nlpsi.copy()
    chi_prev, nlchi_prev = chi.copy(), nlchi.copy()
    psi_rich, chi_rich = psi.copy(), chi.copy()
    advection_rhs(psi_rich, chi_rich, nlpsi, nlchi, uz, work)
    ms.febe(psi_rich, nlpsi, dt)
    ms.febe(chi_rich, nlchi, dt)
    for _ in range(2):
        advection_rhs(psi, chi, nlpsi, nlchi, uz, work)
        ms.febe(psi, nlpsi, dt / 2.0)
        ms.febe(chi, nlchi, dt / 2.0)
    ms.axpby(psi, -1.0 / 1.0, psi_rich, 2.0 / 1.0)    # psi%e = 2*psi%e - psi_rich%e
    ms.axpby(chi, -1.0 / 1.0, chi_rich, 2.0 / 1.0)
    st = VortexState(psi, chi, nlpsi, nlchi, psi_prev, chi_prev, nlpsi_prev, nlchi_prev, uz, work)
    ms.dealias(psi)
    ms.dealias(chi)
    st.gain_psi = ms.svv_filter(psi, st.gain_psi)
    st.gain_chi = ms.svv_filter(chi, st.gain_chi)
    ms.zeroat1(psi)
    ms.zeroat1(chi)
    advection_rhs(psi, chi, nlpsi, nlchi, uz, work)
    return st


def step(st: VortexState, dt: float, check: bool = True):
    """One ABCN step of the main loop; apps/vortical_flow_3d.f90:160-180."""
    ms.abcn(st.psi, st.psi_prev, st.nlpsi, st.nlpsi_prev, dt)
    ms.abcn(st.chi, st.chi_prev, st.nlchi, st.nlchi_prev, dt)
    ms.dealias(st.psi)
    ms.dealias(st.chi)
    st.gain_psi = ms.svv_filter(st.psi, st.gain_psi)
    st.gain_chi = ms.svv_filter(st.chi, st.gain_chi)
    ms.zeroat1(st.psi)
    ms.zeroat1(st.chi)
    advection_rhs(st.psi, st.chi, st.nlpsi, st.nlchi, st.uz, st.work)
    if check:   # check_stability, :397-409: allreduce(land) of the per-rank flags
        from . import dist
        bad = 0.0 if (ms.is_finite(st.psi) and ms.is_finite(st.chi)) else 1.0
        if dist.allreduce([bad])[0] > 0.0:
            raise FloatingPointError("ERROR: non-finite vortex state")
