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


def qvort_dist_tp(kit: TfmKit, q: float = 1.0):
    """apps/vortical_flow_3d.f90:258-326 with ran_noise = 0 (gfortran's random stream is not reproducible)."""
    p = kit.params
    nr, nph, nz = p.nr, p.np // 2, p.nz
    pang = np.array([2.0 * math.acos(-1.0) / p.np * i for i in range(p.np + 1)])
    out = []
    for which in (0, 1):
        glb = np.zeros(kit.glb_sz, dtype=np.complex128, order="F")
        rr_ = kit.r[:, None]
        acc = np.zeros((nr, nph), dtype=np.complex128)
        for xo in (-2, 2):
            yo = 0
            pr = pang[0: 2 * nph: 2][None, :]
            pi_ = pang[1: 2 * nph: 2][None, :]
            rr = np.sqrt((rr_ * np.cos(pr) - xo) ** 2.0 + (rr_ * np.sin(pr) - yo) ** 2.0)
            ri = np.sqrt((rr_ * np.cos(pi_) - xo) ** 2.0 + (rr_ * np.sin(pi_) - yo) ** 2.0)
            den = (1.0 - kit.x[:, None]) ** 2.0
            if which == 0:
                acc = acc + (-np.exp(-(rr ** 2.0)) * 2.0 / den + 1j * (-np.exp(-(ri ** 2.0)) * 2.0 / den))
            else:
                acc = acc + (-np.exp(-(rr ** 2.0)) / q / den + 1j * (-np.exp(-(ri ** 2.0)) / q / den))
        glb[:nr, :nph, :nz] = acc[:, :, None]
        s = Scalar("PPP").upload_global(glb)
        ms.trans(s, "FFF")
        ms.idelsqp(s)
        ms.zeroat1(s)
        out.append(s)
    return out[0], out[1]


def uniform_z_fld(kit: TfmKit, b: float = -0.5) -> Scalar:
    """apps/vortical_flow_3d.f90:328-351."""
    p = kit.params
    glb = np.zeros(kit.glb_sz, dtype=np.complex128, order="F")
    glb[: p.nr, : p.np // 2, : p.nz] = complex(b, b)
    return Scalar("PPP").upload_global(glb)


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
