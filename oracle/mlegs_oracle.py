"""NumPy restatement of the MLegS (v1.1.3) spectral-transform / nonlinear-term path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED (the
reference cannot be compiled here; the analytic known-answers of SURVEY.md
section 4.3, tests/test_oracle_analytic.py, and third-party evaluations of the
tables, tests/test_oracle_independent.py, are what pins this file).

Single process, global arrays.  A field is a complex128 array of shape
(nrdim, npdim, nzdim) in Fortran (column-major) order, i.e. exactly the memory
image of the reference's ``s%e`` on one rank.  Indices below are 0-based; the
reference is 1-based.  Every function cites the reference lines it follows
(paths relative to /root/reference/src; ``ops`` = submodules/mlegs_scalar_ops.f90,
``sinit`` = submodules/mlegs_spectfm_init.f90, ``sdiff`` =
submodules/mlegs_spectfm_diff.f90, ``bops`` = submodules/mlegs_bndmat_ops.f90).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Optional

import numpy as np

PI = math.acos(-1.0)  # modules/mlegs_envir.f90:17


# --------------------------------------------------------------------------- #
# parameters (modules/mlegs_base.f90) and the transform kit
# --------------------------------------------------------------------------- #
@dataclass
class Params:
    nr: int
    np: int
    nz: int
    nrchop: int
    npchop: int
    nzchop: int
    ell: float = 4.0
    zlen: float = 2.0 * PI
    visc: float = 1.0e-3
    hyperpow: int = 0
    hypervisc: float = 0.0
    is_svv: bool = True
    svv_cutoff: float = 0.75
    svv_target: float = 2.0e-2
    svv_strength: float = 0.12
    svv_relax: float = 0.25


def gauss_legendre(n: int):
    """Gauss-Legendre nodes/weights on [-1,1]; sinit:181-214 (Newton, eps 1e-15).

    Vectorised over the node index; every node runs the same recurrence and the
    same stopping test as the scalar loop, so values are identical.
    """
    eps = 1.0e-15
    m = int(math.ceil((n + 1) / 2.0))
    i = np.arange(1, m + 1, dtype=np.float64)
    z = np.cos(PI * (i - 0.25) / (n + 0.5))
    z1 = z + 1.0
    pp = np.zeros_like(z)
    active = np.abs(z - z1) > eps
    while active.any():
        za = z[active]
        p1 = np.ones_like(za)
        p2 = np.zeros_like(za)
        for j in range(1, n + 1):
            p3 = p2
            p2 = p1
            p1 = ((2 * j - 1) * za * p2 - (j - 1) * p3) / j
        ppa = n * (za * p1 - p2) / (za * za - 1.0)
        z1[active] = za
        z[active] = za - p1 / ppa
        pp[active] = ppa
        active = np.abs(z - z1) > eps
    x = np.zeros(n)
    w = np.zeros(n)
    for ii in range(m):
        x[ii] = 0.0 - 1.0 * z[ii]
        x[n - 1 - ii] = 0.0 + 1.0 * z[ii]
        w[ii] = 2.0 * 1.0 / ((1.0 - z[ii] * z[ii]) * pp[ii] * pp[ii])
        w[n - 1 - ii] = w[ii]
    return x, w


def leg_lognorm(ndim: int, ms) -> np.ndarray:
    """log of the normalisation factors, double precision; sinit:218-250."""
    ms = [abs(int(v)) for v in ms]
    me = max(ms)
    wk = np.zeros(me + 1)
    wk[0] = math.log(0.5)
    for m in range(1, me + 1):
        wk[m] = wk[m - 1] + math.log(2.0 * m + 1.0) - math.log(2.0 * m * (2.0 * m - 1.0) ** 2)
    out = np.zeros((ndim, len(ms)), order="F")
    for mm, m in enumerate(ms):
        out[0, mm] = wk[m]
        for nn in range(1, ndim):
            n = m + nn
            out[nn, mm] = (out[nn - 1, mm] + math.log(2.0 * n + 1.0) - math.log(2.0 * n - 1.0)
                           + math.log(1.0 * (n - m)) - math.log(1.0 * (n + m)))
        out[:, mm] = 0.5 * out[:, mm]
    return out


def _libm_lgamma():
    """gfortran's log_gamma is glibc's lgamma; CPython's math.lgamma is its own Lanczos code and
    differs by a few ulp, which exp() turns into ~1e-13 relative table differences."""
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    fn = libm.lgamma
    fn.restype = ctypes.c_double
    fn.argtypes = [ctypes.c_double]
    return fn


_lgamma = _libm_lgamma()


def log_fact(m: int) -> float:
    """sinit:351-357."""
    return _lgamma(2 * m + 1.0) - m * math.log(2.0) - _lgamma(m + 1.0)


def leg_tbl(x, ne: int, ms, lnrm: np.ndarray, digits: int = 50) -> np.ndarray:
    """Normalised associated Legendre table; sinit:254-300.

    The reference runs the three-term recurrence in FM 1.4 at 50 significant
    digits (external/fm1.4/fm_parallel.f90:60) and rounds to double at the end;
    mpmath at ``digits`` plays FM's role here.  Pure-Python loops: small sizes.
    """
    import mpmath as mp

    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    tbl = np.zeros((x.size, ne, len(ms)), order="F")
    with mp.workdps(digits):
        one = mp.mpf(1)
        for mm, mval in enumerate(ms):
            m = abs(int(mval))
            scale = [mp.exp(mp.mpf(log_fact(m + nn)) + mp.mpf(float(lnrm[nn, mm]))) for nn in range(ne)]
            for xx in range(x.size):
                xv = mp.mpf(float(x[xx]))
                col = [None] * ne
                col[0] = (mp.mpf(-1) ** m) * mp.sqrt(one - xv ** 2) ** m
                if ne > 1:
                    col[1] = xv * col[0]
                for nn in range(2, ne):
                    n = m + nn
                    col[nn] = one / (n - m) * (col[nn - 1] * xv - col[nn - 2] * (n + m - 1) / (2 * n - 1) / (2 * n - 3))
                for nn in range(ne):
                    tbl[xx, nn, mm] = float(col[nn] * scale[nn])
    return tbl


@dataclass
class Kit:
    """tfm_kit_3d (modules/mlegs_spectfm.f90:14-63) -- the global ``tfm`` is always 3d."""
    p: Params
    x: np.ndarray
    w: np.ndarray
    ln: np.ndarray
    r: np.ndarray
    lognorm: np.ndarray   # (nrchop+14, npchop)
    pf: np.ndarray        # (nr/2, nrchop+14, npchop)
    at0: np.ndarray
    at1: np.ndarray
    ak: np.ndarray
    m: np.ndarray
    chops: np.ndarray
    nrdim: int
    npdim: int
    nzdim: int
    chopp: int
    chopzl: int
    chopzu: int

    @property
    def glb_sz(self):
        return (self.nrdim, self.npdim, self.nzdim)


def kit_validate(p: Params):
    """sinit:35-62 -- raises ValueError with the reference's stop strings."""
    def smooth235(n):
        while n % 2 == 0 or n % 3 == 0 or n % 5 == 0:
            if n % 2 == 0:
                n //= 2
            if n % 3 == 0:
                n //= 3
            if n % 5 == 0:
                n //= 5
        return n == 1
    if not (p.nr > 0 and p.nr % 2 == 0):
        raise ValueError("tfm_kit_init: nr must be even")
    if not (p.np > 0 and (p.np == 1 or p.np % 2 == 0)):
        raise ValueError("tfm_kit_init: np must be even")
    if p.np != 1 and not smooth235(p.np):
        raise ValueError("tfm_kit_init: np must only have factors of 2, 3 and 5")
    if not (p.nz > 0 and (p.nz == 1 or p.nz % 2 == 0)):
        raise ValueError("tfm_kit_init: nz must be even")
    if p.nz != 1 and not smooth235(p.nz):
        raise ValueError("tfm_kit_init: nz must only have factors of 2, 3 and 5")
    if p.nrchop > p.nr:
        raise ValueError("tfm_kit_init: nrchop must be smaller than or equal to nr")
    if p.np != 1 and p.npchop * 2 > p.np + 2:
        raise ValueError("tfm_kit_init: npchop <= np/2 + 1 must be satisfied")
    if p.nz != 1 and p.nzchop * 2 > p.nz + 2:
        raise ValueError("tfm_kit_init: nzchop <= nz/2 + 1 must be satisfied")


def kit_init(p: Params, tables: Optional[dict] = None) -> Kit:
    """tfm_kit_init for the 3d kit; sinit:6-154 (110-150).

    ``tables`` (keys x,w,lognorm,pf,at0,at1) lets a caller inject tables built
    elsewhere -- the same arrays the C-ABI ``mlegs_b200_init`` receives -- so
    oracle and device share bit-identical tables (SURVEY.md section 7 "Tables").
    """
    kit_validate(p)
    nr, nrchop, npchop = p.nr, p.nrchop, p.npchop
    if tables is None:
        x, w = gauss_legendre(nr)
    else:
        x, w = np.asarray(tables["x"], dtype=np.float64), np.asarray(tables["w"], dtype=np.float64)
    ln = np.array([-math.log(1.0 - float(v)) for v in x])      # libm log, like gfortran
    r = p.ell * np.sqrt((1.0 + x) / (1.0 - x))
    chops = np.array([max(nrchop - i, 0) for i in range(npchop)], dtype=np.int64)
    nrdim = nr + max(3, p.hyperpow)
    m = np.arange(npchop, dtype=np.int64)
    npdim = p.np // 2 + 1
    ak = np.array([2.0 * PI / p.zlen * i for i in range(-p.nz, 0)])      # sinit:133
    ak[: p.nz // 2 + 1] = np.array([2.0 * PI / p.zlen * i for i in range(0, p.nz // 2 + 1)])  # sinit:134
    nzdim = p.nz
    if tables is None:
        lognorm = leg_lognorm(nrchop + 14, m)
        pf = leg_tbl(x[: nr // 2], nrchop + 14, m, lognorm)
        at0 = leg_tbl([-1.0 + 1.0e-15], nrchop, [0], lognorm[:nrchop, :1]).reshape(nrchop)
        at1 = leg_tbl([1.0 - 1.0e-15], nrchop, [0], lognorm[:nrchop, :1]).reshape(nrchop)
    else:
        lognorm = np.asfortranarray(tables["lognorm"], dtype=np.float64)
        pf = np.asfortranarray(tables["pf"], dtype=np.float64)
        at0 = np.asarray(tables["at0"], dtype=np.float64)
        at1 = np.asarray(tables["at1"], dtype=np.float64)
    return Kit(p=p, x=x, w=w, ln=ln, r=r, lognorm=lognorm, pf=pf, at0=at0, at1=at1, ak=ak, m=m,
               chops=chops, nrdim=nrdim, npdim=npdim, nzdim=nzdim, chopp=npchop,
               chopzl=p.nzchop, chopzu=p.nz - p.nzchop + 2)


# --------------------------------------------------------------------------- #
# the scalar (modules/mlegs_scalar.f90:15-49), single rank: loc == glb
# --------------------------------------------------------------------------- #
@dataclass
class Scalar:
    e: np.ndarray
    space: str = "PPP"
    ln: float = 0.0
    nrchop_offset: int = 0
    npchop_offset: int = 0
    nzchop_offset: int = 0

    def copy(self) -> "Scalar":     # scalar_copy, submodules/mlegs_scalar_init.f90:106-140
        return replace(self, e=self.e.copy(order="F"))

    def chop_offset(self, iof1, iof2=0, iof3=0):   # mlegs_scalar_init.f90:84-103
        self.nrchop_offset, self.npchop_offset, self.nzchop_offset = iof1, iof2, iof3


def scalar_init(kit: Kit, space: str = "PPP") -> Scalar:
    return Scalar(e=np.zeros(kit.glb_sz, dtype=np.complex128, order="F"), space=space)


@dataclass
class ChopIdx:
    nrc: int
    npc: int
    nzc: int
    nzcu: int
    nrcs: np.ndarray


def chop_index(s: Scalar, kit: Kit) -> ChopIdx:
    """ops:2012-2063."""
    nrc = int(kit.chops[0]) + s.nrchop_offset
    npc = kit.chopp + s.npchop_offset
    nzc = kit.chopzl + s.nzchop_offset
    nzcu = kit.chopzu - s.nzchop_offset
    if nrc > kit.nrdim:
        raise ValueError("chop_index: chopping in r too large")
    if npc > kit.npdim:
        raise ValueError("chop_index: chopping in p too large")
    if 2 * s.nzchop_offset > nzcu - nzc:
        raise ValueError("chop_index: chopping in z too large")
    nrcs = np.zeros(max(npc, kit.chopp), dtype=np.int64)
    for mm in range(npc):
        nrcs[mm] = max(min(nrc, nrc - mm), 0)
    return ChopIdx(nrc, npc, nzc, nzcu, nrcs)


def k_ranges(ci: ChopIdx, nzdim: int):
    """The two axial loops every banded operator runs (e.g. ops:546-557):
    kk = 1..nzc, then kk = max(nzcu,1)..nzdim (1-based).  When
    nzchop == nz/2+1 the Nyquist plane is in BOTH ranges and is processed twice;
    that is the reference's behaviour and is kept."""
    first = list(range(0, min(nzdim, ci.nzc)))
    second = list(range(max(ci.nzcu, 1) - 1, nzdim))
    return first, second


# --------------------------------------------------------------------------- #
# masks and filters
# --------------------------------------------------------------------------- #
def chop(s: Scalar, kit: Kit):
    """ops:6-41."""
    ci = chop_index(s, kit)
    e = s.e
    if s.space[0] == "F":
        for mm in range(e.shape[1]):
            rc = int(ci.nrcs[mm]) if mm < ci.nrcs.size else 0
            e[rc:, mm, :] = 0.0
    if s.space[1] == "F":
        e[:, ci.npc:, :] = 0.0
    if s.space[2] == "F":
        if ci.nzc < ci.nzcu:
            e[:, :, ci.nzc: ci.nzcu - 1] = 0.0


def dealias(s: Scalar, kit: Kit):
    """2/3 rule in the periodic directions; ops:43-70."""
    p = kit.p
    if p.np > 1 and s.space[1] == "F":
        pcut = max(p.np // 3 + 1, 1)
        s.e[:, pcut:, :] = 0.0            # global_index(1-based) > pcut
    if p.nz > 1 and s.space[2] == "F":
        zcut = max(p.nz // 3 + 1, 1)
        zupper = p.nz - zcut + 2
        # 1-based zcut < g < zupper  ->  0-based zcut .. zupper-2
        s.e[:, :, zcut: zupper - 1] = 0.0


def svv_filter(s: Scalar, kit: Kit, gain: float) -> float:
    """Tail-energy driven spectral vanishing viscosity; ops:72-155.  Returns the new gain."""
    p = kit.p
    if not p.is_svv:
        return gain
    if s.space != "FFF":
        raise ValueError("svv_filter: scalar must be in FFF space")
    e = s.e
    total = float(np.sum(np.abs(e) ** 2))
    cutoff = min(max(p.svv_cutoff, 0.0), 0.99)
    target = max(p.svv_target, 1.0e-12)
    kmax = max(float(np.max(np.abs(kit.ak))), 1.0)
    n1, n2, n3 = e.shape
    q_r = np.minimum(1.0, np.maximum(np.arange(n1), 0).astype(np.float64) / max(float(p.nrchop - 1), 1.0))
    q_p = np.minimum(1.0, np.abs(np.arange(n2)).astype(np.float64) / max(float(p.np // 2), 1.0))
    kval = np.zeros(n3)
    kval[: min(n3, kit.ak.size)] = np.abs(kit.ak[: min(n3, kit.ak.size)])
    q_z = np.minimum(1.0, kval / kmax)
    q = np.minimum(1.0, np.maximum(np.maximum(q_r[:, None, None], q_p[None, :, None]), q_z[None, None, :]))
    tail = float(np.sum((np.abs(e) ** 2)[q >= cutoff]))
    if total <= np.finfo(np.float64).tiny:
        return gain
    tail_ratio = tail / total
    feedback = min(max(tail_ratio / target - 1.0, 0.0), 1.0)
    relax = min(max(p.svv_relax, 0.0), 1.0)
    gain = (1.0 - relax) * gain + relax * feedback
    strength = min(max(p.svv_strength, 0.0) * gain, 1.0)
    if strength <= 0.0:
        return gain
    msk = q > cutoff
    shape = (q - cutoff) / (1.0 - cutoff)
    factor = np.exp(-strength * shape ** 8)
    e[msk] = factor[msk] * e[msk]
    return gain


# --------------------------------------------------------------------------- #
# transforms
# --------------------------------------------------------------------------- #
def horizontal_fft_forward(s: Scalar, kit: Kit):
    """ops:1567-1641.  FFTE dzfft2d(NY=1) == rfft (external/ffte-7.0/dzfft2d.f:96-103), then /np."""
    npts = kit.p.np
    nph = npts // 2
    e = s.e
    c = np.empty((e.shape[0], npts, e.shape[2]))
    c[:, 0::2, :] = e[:, :nph, :].real
    c[:, 1::2, :] = e[:, :nph, :].imag
    b = np.fft.rfft(c, axis=1)
    e[:, : nph + 1, :] = b / (2 * nph)


def horizontal_fft_backward(s: Scalar, kit: Kit):
    """ops:1645-1717.  zdfft2d(NY=1) is a Hermitian inverse that ignores Im(m=0) and
    Im(Nyquist) (external/ffte-7.0/zdfft2d.f:119-128); the 1/N is undone by *(2*nph).
    The padding column keeps the Nyquist input times np (quirk Q3)."""
    npts = kit.p.np
    nph = npts // 2
    e = s.e
    cin = e[:, : nph + 1, :].copy()
    d = np.fft.irfft(cin, n=npts, axis=1) * npts
    e[:, :nph, :] = d[:, 0::2, :] + 1j * d[:, 1::2, :]
    e[:, nph, :] = cin[:, nph, :] * (2 * nph)


def vertical_fft_forward(s: Scalar, kit: Kit):
    """ops:1721-1783: zfft1d forward (e^{-i}), then /nz."""
    nz = kit.p.nz
    s.e[:, :, :nz] = np.fft.fft(s.e[:, :, :nz], axis=2) / nz


def vertical_fft_backward(s: Scalar, kit: Kit):
    """ops:1787-1848: zfft1d inverse (conj-fwd-conj/N, external/ffte-7.0/zfft1d.f:41-47,76-83), then *nz."""
    nz = kit.p.nz
    s.e[:, :, :nz] = np.fft.ifft(s.e[:, :, :nz], axis=2) * nz


def _rtrans_nrcs(s: Scalar, kit: Kit):
    nrc = int(kit.chops[0]) + s.nrchop_offset
    npc = kit.chopp + s.npchop_offset
    if nrc > kit.nrdim:
        raise ValueError("rtrans: chopping in r too large")
    if npc > kit.npdim:
        raise ValueError("rtrans: chopping in p too large")
    return npc, [max(min(nrc, nrc - mm), 0) for mm in range(npc)]


def rtrans_forward(s: Scalar, kit: Kit):
    """Even/odd folded mapped-Legendre analysis; ops:1852-1927."""
    nr = kit.p.nr
    nrh = nr // 2
    npc, nrcs = _rtrans_nrcs(s, kit)
    e = s.e
    se = np.zeros_like(e)
    w = kit.w[:nrh, None]
    for mm in range(min(e.shape[1], npc)):
        top = e[:nrh, mm, :]
        bot = e[nr - 1: nrh - 1: -1, mm, :]     # e(nr-i+1), i=1..nrh
        be = (top + bot) * w
        bo = (top - bot) * w
        nn = nrcs[mm]
        if nn >= 1:
            se[0:nn:2, mm, :] = kit.pf[:nrh, 0:nn:2, mm].T @ be
        if nn >= 2:
            se[1:nn:2, mm, :] = kit.pf[:nrh, 1:nn:2, mm].T @ bo
    s.e = np.asfortranarray(se)


def rtrans_backward(s: Scalar, kit: Kit):
    """Even/odd folded mapped-Legendre synthesis; ops:1931-2008."""
    nr = kit.p.nr
    nrh = nr // 2
    npc, nrcs = _rtrans_nrcs(s, kit)
    e = s.e
    se = np.zeros_like(e)
    for mm in range(min(e.shape[1], npc)):
        nn = nrcs[mm]
        be = kit.pf[:nrh, 0:nn:2, mm] @ e[0:nn:2, mm, :] if nn >= 1 else np.zeros((nrh, e.shape[2]), complex)
        bo = kit.pf[:nrh, 1:nn:2, mm] @ e[1:nn:2, mm, :] if nn >= 2 else np.zeros((nrh, e.shape[2]), complex)
        se[:nrh, mm, :] = be + bo
        se[nr - 1: nrh - 1: -1, mm, :] = be - bo
    s.e = np.asfortranarray(se)


_SPACES = {"PPP": 0, "PFP": 1, "FFP": 2, "FFF": 3}


def trans(s: Scalar, space: str, kit: Kit):
    """State machine PPP <-> PFP <-> FFP <-> FFF; ops:157-235 (exchanges are no-ops on one rank)."""
    if s.space not in _SPACES:
        raise ValueError("trans: scalar space info corrupted (only accepting PPP, PFP, FFP and FFF)")
    if space not in _SPACES:
        raise ValueError("trans: only taking PPP, PFP, FFP and FFF for spectral transformation")
    p = kit.p
    cur, new = _SPACES[s.space], _SPACES[space]
    nr = p.nr
    while cur < new:
        if cur == 0:
            if p.np > 1:
                horizontal_fft_forward(s, kit)
            s.space = "PFP"
        elif cur == 1:
            s.e[:nr, 0, :] = s.e[:nr, 0, :] - (s.ln * kit.ln[:nr])[:, None]
            rtrans_forward(s, kit)
            s.space = "FFP"
        elif cur == 2:
            if p.nz > 1:
                vertical_fft_forward(s, kit)
            s.space = "FFF"
        cur += 1
    while cur > new:
        if cur == 3:
            if p.nz > 1:
                vertical_fft_backward(s, kit)
            s.space = "FFP"
        elif cur == 2:
            rtrans_backward(s, kit)
            s.e[:nr, 0, :] = s.e[:nr, 0, :] + (s.ln * kit.ln[:nr])[:, None]
            s.space = "PFP"
        elif cur == 1:
            if p.np > 1:
                horizontal_fft_backward(s, kit)
            s.space = "PPP"
        cur -= 1


def _calcat(s: Scalar, kit: Kit, at: np.ndarray) -> np.ndarray:
    ci = chop_index(s, kit)
    st = s.copy()
    trans(st, "FFF", kit)
    nrc = min(st.e.shape[0], ci.nrc)
    return st.e[:nrc, 0, :].T @ at[:nrc]


def calcat0(s: Scalar, kit: Kit) -> np.ndarray:
    """value at r = 0 per axial mode; ops:237-272."""
    return _calcat(s, kit, kit.at0)


def calcat1(s: Scalar, kit: Kit) -> np.ndarray:
    """value at r -> inf per axial mode; ops:274-309."""
    return _calcat(s, kit, kit.at1)


def zeroat1(s: Scalar, kit: Kit):
    """ops:311-325."""
    if s.space != "FFF":
        trans(s, "FFF", kit)
    at1 = calcat1(s, kit)
    s.e[0, 0, :] = s.e[0, 0, :] - at1 / kit.at1[0]


# --------------------------------------------------------------------------- #
# banded operators.  A band matrix is a dict {d: vector}: A[i, i+d] = band[d][i]
# --------------------------------------------------------------------------- #
_SCALE_CACHE: dict = {}


def _scale_factors(lognorm_col: np.ndarray, d: int) -> np.ndarray:
    """exp(lognorm(i+d) - lognorm(i)) through libm's exp (what gfortran calls; numpy's SIMD exp can
    differ by an ulp), cached per lognorm column and diagonal."""
    key = (hash(lognorm_col.tobytes()), lognorm_col.size, d)
    fac = _SCALE_CACHE.get(key)
    if fac is None:
        nmax = lognorm_col.size
        fac = np.ones(nmax)
        for i in range(nmax):
            j = i + d
            if 0 <= j < nmax:
                fac[i] = math.exp(lognorm_col[j] - lognorm_col[i])
        _SCALE_CACHE[key] = fac
    return fac


def _scale_band(band: dict, lognorm_col: np.ndarray, n: int) -> dict:
    """genm(i,j) *= exp(lognorm(j) - lognorm(i)); e.g. sdiff:146-150."""
    out = {}
    i = np.arange(n)
    for d, v in band.items():
        j = i + d
        ok = (j >= 0) & (j < n)
        fac = _scale_factors(lognorm_col, d)[:n]
        out[d] = np.where(ok, v * fac, 0.0)
    return out


def leg_xxdx(mval: int, n: int, kit: Kit) -> dict:
    """(1-x)^2 d/dx == r d/dr, tridiagonal; sdiff:6-43 (n_input == n_output == n)."""
    am = abs(mval)
    nn = np.arange(n)
    nv = (am + nn).astype(np.float64)
    sub = -(nv - 1.0) * (nv - am) / (2.0 * nv - 1.0)
    sup = (nv + 2.0) * (nv + am + 1.0) / (2.0 * nv + 3.0)
    band = {-1: np.where(nn - 1 >= 0, sub, 0.0), 1: np.where(nn + 1 < n, sup, 0.0)}
    return _scale_band(band, kit.lognorm[:, am], n)


def _del2_raw(mval: int, n: int, ell: float):
    am = abs(mval)
    s = 0.0
    nn = np.arange(n)
    nv = (am + nn).astype(np.float64)
    ell2 = ell ** 2.0
    dm2 = -(nv - am - 1.0) * (nv - am) * (nv - 2.0 + s) * (nv - 1.0 + s) / (2.0 * nv - 3.0) / (2.0 * nv - 1.0)
    dm1 = 2.0 * nv * (nv - am) * (nv - 1.0 + s) / (2.0 * nv - 1.0)
    d0 = (-2.0 * nv * (nv + 1.0) * (3.0 * nv * nv + 3.0 * nv - am * am - 2.0)
          + 2.0 * s * (s - 2.0) * (nv * nv + nv + am * am - 1.0)) / (2.0 * nv - 1.0) / (2.0 * nv + 3.0)
    dp1 = 2.0 * (nv + 1.0) * (nv + am + 1.0) * (nv + 2.0 - s) / (2.0 * nv + 3.0)
    dp2 = -(nv + am + 1.0) * (nv + am + 2.0) * (nv + 3.0 - s) * (nv + 2.0 - s) / (2.0 * nv + 3.0) / (2.0 * nv + 5.0)
    band = {-2: np.where(nn - 2 >= 0, dm2, 0.0), -1: np.where(nn - 1 >= 0, dm1, 0.0), 0: d0,
            1: np.where(nn + 1 < n, dp1, 0.0), 2: np.where(nn + 2 < n, dp2, 0.0)}
    return {d: v / ell2 for d, v in band.items()}


def leg_del2h(mval: int, n: int, kit: Kit) -> dict:
    """horizontal Laplacian, pentadiagonal; sdiff:45-96."""
    return _scale_band(_del2_raw(mval, n, kit.p.ell), kit.lognorm[:, abs(mval)], n)


def leg_del2(mval: int, akval: float, n: int, kit: Kit) -> dict:
    """Laplacian, pentadiagonal, -ak^2 on the diagonal; sdiff:98-152."""
    band = _del2_raw(mval, n, kit.p.ell)
    band[0] = band[0] - akval ** 2.0
    return _scale_band(band, kit.lognorm[:, abs(mval)], n)


def band_mulvec(band: dict, v: np.ndarray) -> np.ndarray:
    """band .mul. vector(s): out(j) = sum_k A(j,k) v(k), k ascending (bops:123-133 via 92-117).
    ``v`` may be (n,) or (n, ncols)."""
    n = v.shape[0]
    out = np.zeros_like(v, dtype=np.complex128)
    for d in sorted(band):
        a = band[d]
        if d >= 0:
            rows = slice(0, n - d)
            cols = slice(d, n)
        else:
            rows = slice(-d, n)
            cols = slice(0, n + d)
        if v.ndim == 1:
            out[rows] = out[rows] + a[rows] * v[cols]
        else:
            out[rows] = out[rows] + a[rows, None] * v[cols]
    return out


def band_mulband(a: dict, b: dict, n: int) -> dict:
    """band .mul. band for n x n matrices, C(r,j) = sum_k A(r,k) B(k,j) with k ascending;
    bops:251-318 (mulrbrb)."""
    out = {}
    r = np.arange(n)
    for dc in range(min(a) + min(b), max(a) + max(b) + 1):
        acc = np.zeros(n)
        for da in sorted(a):
            db = dc - da
            if db not in b:
                continue
            k = r + da
            j = r + dc
            ok = (k >= 0) & (k < n) & (j >= 0) & (j < n)
            kk = np.clip(k, 0, n - 1)
            acc = acc + np.where(ok, a[da] * b[db][kk], 0.0)
        out[dc] = acc
    return out


def band_to_lapack(band: dict, n: int, kl: int, ku: int) -> np.ndarray:
    """ab(kl+ku+1+i-j, j) = A(i,j) in the (2kl+ku+1, n) layout zgbtrf wants (bops:388-395)."""
    ab = np.zeros((2 * kl + ku + 1, n), dtype=np.complex128, order="F")
    for d, v in band.items():
        if d > ku or -d > kl:
            continue
        i = np.arange(n)
        j = i + d
        ok = (j >= 0) & (j < n)
        ab[kl + ku + i[ok] - j[ok], j[ok]] = v[ok]
    return ab


def band_lsolve(band: dict, n: int, kl: int, ku: int, b: np.ndarray) -> np.ndarray:
    """lsolve == zgbtrf + zgbtrs on the real matrix promoted to complex (bops:441-457,406-430)."""
    from scipy.linalg import lapack
    ab = band_to_lapack(band, n, kl, ku)
    lu, piv, info = lapack.zgbtrf(ab, kl, ku)
    if info != 0:
        raise ValueError("lurc: lu factorization resulted in failure")
    x, info = lapack.zgbtrs(lu, kl, ku, np.asarray(b, dtype=np.complex128), piv)
    if info != 0:
        raise ValueError("solvecbc: linear system unable to be solved")
    return x


def _require_fff(s: Scalar, kit: Kit):
    if s.space != "FFF":
        trans(s, "FFF", kit)


def delsqp(s: Scalar, kit: Kit):
    """(1-x)^{-2} del^2_perp, diagonal; ops:327-366."""
    ci = chop_index(s, kit)
    _require_fff(s, kit)
    so = s.copy()
    so.ln = 0.0
    ell = kit.p.ell
    for mm in range(min(so.e.shape[1], ci.npc)):
        for nn in range(min(so.e.shape[0], int(ci.nrcs[mm]))):
            n = int(kit.m[mm]) + nn
            so.e[nn, mm, :] = -s.e[nn, mm, :] * n * (n + 1.0) / (ell ** 2.0)
    so.e[0, 0, 0] = s.ln / ell ** 2.0 / math.exp(kit.lognorm[0, 0])
    s.e, s.ln = so.e, so.ln


def idelsqp(s: Scalar, kit: Kit):
    """inverse of delsqp; ops:368-416."""
    ci = chop_index(s, kit)
    _require_fff(s, kit)
    so = s.copy()
    ell = kit.p.ell
    so.ln = float((s.e[0, 0, 0] * ell ** 2.0 * math.exp(kit.lognorm[0, 0])).real)
    for mm in range(min(so.e.shape[1], ci.npc)):
        for nn in range(min(so.e.shape[0], int(ci.nrcs[mm]))):
            n = int(kit.m[mm]) + nn
            if n == 0:
                so.e[nn, mm, :] = 0.0
            else:
                so.e[nn, mm, :] = -s.e[nn, mm, :] / n / (n + 1.0) * (ell ** 2.0)
    so.e[0, 0, :] = 0.0
    s.e, s.ln = so.e, so.ln


# The per-(m,k) sweeps below (ops:442-449, 544-558, 600-668, 700-758, 828-852, 953-984) are independent per azimuthal
# column m.  set_workers(n > 1) runs the columns of one sweep on n forked processes -- the same function on the same
# column data, so results are bit-identical to the serial sweep (tests/test_oracle_analytic.py checks that); it only
# makes the 64^3 / 128^3 time-step parity tests affordable.  Never call it in a process that has initialised CUDA.
_WORKERS = 1
_SWEEP = None


def set_workers(n: int):
    global _WORKERS
    _WORKERS = max(1, int(n))


def _sweep_task(mm: int):
    body, e = _SWEEP
    col = np.array(e[:, mm, :], order="F")
    aux = body(mm, col)
    return mm, col, aux


def _sweep_columns(s: Scalar, cols, body) -> dict:
    """body(mm, col) updates col = s.e[:, mm, :] in place and returns an auxiliary value (or None).  Returns {mm: aux}."""
    global _SWEEP
    cols = list(cols)
    if _WORKERS <= 1 or len(cols) < 2:
        return {mm: body(mm, s.e[:, mm, :]) for mm in cols}
    import multiprocessing as mp
    _SWEEP = (body, s.e)
    try:
        with mp.get_context("fork").Pool(min(_WORKERS, len(cols))) as pool:
            out = pool.map(_sweep_task, cols, chunksize=1)
    finally:
        _SWEEP = None
    aux = {}
    for mm, col, a in out:
        s.e[:, mm, :] = col
        aux[mm] = a
    return aux


def _apply_band_per_mk(s: Scalar, kit: Kit, builder, per_k: bool):
    """Shared loop of xxdx/del2h/del2 (ops:442-449, 489-503, 544-558): the operator acts on rows
    :nn of every retained (m,k) column, first k-range then second k-range."""
    ci = chop_index(s, kit)
    first, second = k_ranges(ci, s.e.shape[2])

    def body(mm, col):
        nn = int(ci.nrcs[mm])
        if nn < 1:
            return None
        mval = int(kit.m[mm])
        if not per_k:
            band = builder(mval, 0.0, nn)
            for rng in (first, second):
                if rng:
                    col[:nn, rng[0]: rng[-1] + 1] = band_mulvec(band, col[:nn, rng[0]: rng[-1] + 1])
        else:
            for rng in (first, second):
                for kk in rng:
                    band = builder(mval, float(kit.ak[kk]), nn)
                    col[:nn, kk] = band_mulvec(band, col[:nn, kk])
        return None

    _sweep_columns(s, range(min(s.e.shape[1], ci.npc)), body)


def xxdx(s: Scalar, kit: Kit):
    """r d/dr in spectral space; ops:418-463."""
    _require_fff(s, kit)
    ln = s.ln
    s.ln = 0.0
    _apply_band_per_mk(s, kit, lambda m, ak, nn: leg_xxdx(m, nn, kit), per_k=False)
    s.e[0, 0, 0] = s.e[0, 0, 0] + 1.0 / math.exp(kit.lognorm[0, 0]) * ln
    s.e[1, 0, 0] = s.e[1, 0, 0] + 1.0 / math.exp(kit.lognorm[1, 0]) * ln


def _ln_del2_terms(kit: Kit):
    ell = kit.p.ell
    return (4.0 / 3.0 / ell ** 2.0 / math.exp(kit.lognorm[0, 0]),
            2.0 / 1.0 / ell ** 2.0 / math.exp(kit.lognorm[1, 0]),
            2.0 / 3.0 / ell ** 2.0 / math.exp(kit.lognorm[2, 0]))


def del2h(s: Scalar, kit: Kit):
    """del^2_perp; ops:465-518."""
    _require_fff(s, kit)
    ln = s.ln
    s.ln = 0.0
    _apply_band_per_mk(s, kit, lambda m, ak, nn: leg_del2h(m, nn, kit), per_k=False)
    c0, c1, c2 = _ln_del2_terms(kit)
    s.e[0, 0, 0] = s.e[0, 0, 0] + c0 * ln
    s.e[1, 0, 0] = s.e[1, 0, 0] - c1 * ln
    s.e[2, 0, 0] = s.e[2, 0, 0] + c2 * ln


def del2(s: Scalar, kit: Kit):
    """del^2; ops:520-573."""
    _require_fff(s, kit)
    ln = s.ln
    s.ln = 0.0
    _apply_band_per_mk(s, kit, lambda m, ak, nn: leg_del2(m, ak, nn, kit), per_k=True)
    c0, c1, c2 = _ln_del2_terms(kit)
    s.e[0, 0, 0] = s.e[0, 0, 0] + c0 * ln
    s.e[1, 0, 0] = s.e[1, 0, 0] - c1 * ln
    s.e[2, 0, 0] = s.e[2, 0, 0] + c2 * ln


def idel2_proln(s: Scalar, kit: Kit):
    """Poisson solve, ln taken from the (1,1,1) entry; ops:671-760."""
    ci = chop_index(s, kit)
    _require_fff(s, kit)
    ell = kit.p.ell
    first, second = k_ranges(ci, s.e.shape[2])

    def body(mm, col):
        ln = None
        nn = int(ci.nrcs[mm])
        if nn < 1:
            return ln
        mval = int(kit.m[mm])
        for ir, rng in enumerate((first, second)):
            for kk in rng:
                band = leg_del2(mval, float(kit.ak[kk]), nn, kit)
                if ir == 0 and mm == 0 and kk == 0:
                    band[0][0] = band[0][0] + 4.0 / 3.0 / ell ** 2.0
                    band[-1][1] = band[-1][1] - 2.0 / 1.0 / ell ** 2.0 * math.exp(kit.lognorm[0, 0] - kit.lognorm[1, 0])
                    band[-2][2] = band[-2][2] + 2.0 / 3.0 / ell ** 2.0 * math.exp(kit.lognorm[0, 0] - kit.lognorm[2, 0])
                    col[:nn, kk] = band_lsolve(band, nn, 2, 2, col[:nn, kk])
                    ln = float((col[0, kk] * math.exp(kit.lognorm[0, 0])).real)
                else:
                    col[:nn, kk] = band_lsolve(band, nn, 2, 2, col[:nn, kk])
        return ln

    aux = _sweep_columns(s, range(min(s.e.shape[1], ci.npc)), body)
    if aux.get(0) is not None:
        s.ln = aux[0]


def idel2_preln(s: Scalar, kit: Kit, preln: float):
    """Poisson solve with a prescribed ln; ops:575-669 (shifted first column, sublen=3)."""
    ci = chop_index(s, kit)
    _require_fff(s, kit)
    ell = kit.p.ell
    first, second = k_ranges(ci, s.e.shape[2])

    def body(mm, col):
        ln = None
        nn = int(ci.nrcs[mm])
        if nn < 1:
            return ln
        mval = int(kit.m[mm])
        for ir, rng in enumerate((first, second)):
            for kk in rng:
                band = leg_del2(mval, float(kit.ak[kk]), nn, kit)
                if ir == 0 and mm == 0 and kk == 0:
                    full = np.zeros((nn, nn))
                    for d, v in band.items():
                        i = np.arange(nn)
                        j = i + d
                        ok = (j >= 0) & (j < nn)
                        full[i[ok], j[ok]] = v[ok]
                    full[1:nn, :] = full[0:nn - 1, :].copy()
                    full[0, :] = 0.0
                    full[0, 0] = 1.0
                    full[1, 0] = full[1, 0] + 4.0 / 3.0 / ell ** 2.0
                    full[2, 0] = full[2, 0] - 2.0 / 1.0 / ell ** 2.0 * math.exp(kit.lognorm[0, 0] - kit.lognorm[1, 0])
                    full[3, 0] = full[3, 0] + 2.0 / 3.0 / ell ** 2.0 * math.exp(kit.lognorm[0, 0] - kit.lognorm[2, 0])
                    b2 = {}
                    for d in range(-3, 3):
                        i = np.arange(nn)
                        j = i + d
                        ok = (j >= 0) & (j < nn)
                        v = np.zeros(nn)
                        v[ok] = full[i[ok], j[ok]]
                        b2[d] = v
                    rhs = col[:nn, kk].copy()
                    rhs[1:nn] = col[0:nn - 1, kk]
                    rhs[0] = preln / math.exp(kit.lognorm[0, 0])
                    col[:nn, kk] = band_lsolve(b2, nn, 3, 2, rhs)
                    ln = float((col[0, kk] * math.exp(kit.lognorm[0, 0])).real)
                else:
                    col[:nn, kk] = band_lsolve(band, nn, 2, 2, col[:nn, kk])
        return ln

    aux = _sweep_columns(s, range(min(s.e.shape[1], ci.npc)), body)
    if aux.get(0) is not None:
        s.ln = aux[0]


def ihelm(s: Scalar, alpha: float, kit: Kit):
    """solve (del^2 + alpha) x = s; ops:791-854."""
    ci = chop_index(s, kit)
    _require_fff(s, kit)
    if abs(alpha) < 5.0e-14:
        raise ValueError("ihelm: alpha equals to zero. Inversion of 0*identity is impossible")
    s.ln = s.ln / alpha
    c0, c1, c2 = _ln_del2_terms(kit)
    s.e[0, 0, 0] = s.e[0, 0, 0] - c0 * s.ln
    s.e[1, 0, 0] = s.e[1, 0, 0] + c1 * s.ln
    s.e[2, 0, 0] = s.e[2, 0, 0] - c2 * s.ln
    first, second = k_ranges(ci, s.e.shape[2])

    def body(mm, col):
        nn = int(ci.nrcs[mm])
        if nn < 1:
            return None
        mval = int(kit.m[mm])
        for rng in (first, second):
            for kk in rng:
                band = leg_del2(mval, float(kit.ak[kk]), nn, kit)
                band[0] = band[0] + alpha
                col[:nn, kk] = band_lsolve(band, nn, 2, 2, col[:nn, kk])
        return None

    _sweep_columns(s, range(min(s.e.shape[1], ci.npc)), body)


def helmp(s: Scalar, power: int, alpha: float, beta: float, kit: Kit):
    """del^p + beta del^2 + alpha; ops:856-903."""
    if not (power % 2 == 0 and power >= 4):
        raise ValueError("helmp: even power greater than or equal to 4")
    if power > 8:
        raise ValueError("helmp: power must be less than or equal to 8 (supported power = 4, 6 or 8)")
    _require_fff(s, kit)
    s2 = s.copy()
    del2(s2, kit)
    sp = s2.copy()
    for _ in range(power // 2 - 1):
        del2(sp, kit)
    s.e = np.asfortranarray(sp.e + beta * s2.e + alpha * s.e)
    s.ln = alpha * s.ln


def helmp_band(mval: int, akval: float, nn: int, power: int, alpha: float, beta: float, kit: Kit) -> dict:
    """The banded matrix ihelmp factors for one (m,k); ops:958-965."""
    d2 = leg_del2(mval, akval, nn, kit)
    hp = d2
    for _ in range(power // 2 - 1):
        hp = band_mulband(d2, hp, nn)
    out = {}
    for d in range(-power, power + 1):
        v = hp.get(d, np.zeros(nn)).copy()
        v = v + beta * d2.get(d, np.zeros(nn))
        out[d] = v
    out[0] = out[0] + alpha
    return out


def ihelmp(s: Scalar, power: int, alpha: float, beta: float, kit: Kit):
    """solve (del^p + beta del^2 + alpha) x = s; ops:905-1000."""
    if not (power % 2 == 0 and power >= 4):
        raise ValueError("ihelmp: even power greater than or equal to 4")
    if power > 8:
        raise ValueError("ihelmp: power must be less than or equal to 8 (supported power = 4, 6 or 8)")
    ci = chop_index(s, kit)
    _require_fff(s, kit)
    if abs(alpha) < 5.0e-14:
        raise ValueError("ihelmp: alpha equals to zero. Inversion of 0*identity is impossible")
    s.ln = s.ln / alpha
    nrc = ci.nrc
    c0, c1, c2 = _ln_del2_terms(kit)
    bl2 = np.zeros(nrc)
    bl2[0] = c0 * s.ln
    bl2[1] = -c1 * s.ln
    bl2[2] = c2 * s.ln
    bl = bl2.copy()
    d2_00 = leg_del2(int(kit.m[0]), float(kit.ak[0]), nrc, kit)
    for _ in range(power // 2 - 1):
        bl = band_mulvec(d2_00, bl.astype(np.complex128)).real
    s.e[:nrc, 0, 0] = s.e[:nrc, 0, 0] - bl - beta * bl2
    first, second = k_ranges(ci, s.e.shape[2])

    def body(mm, col):
        nn = int(ci.nrcs[mm])
        if nn < 1:
            return None
        mval = int(kit.m[mm])
        for rng in (first, second):
            for kk in rng:
                band = helmp_band(mval, float(kit.ak[kk]), nn, power, alpha, beta, kit)
                col[:nn, kk] = band_lsolve(band, nn, power, power, col[:nn, kk])
        return None

    _sweep_columns(s, range(min(s.e.shape[1], ci.npc)), body)


def smooth(a: np.ndarray) -> np.ndarray:
    """ops:2149-2175: weighted three-point smoothing of a far-field tail, the last point spread over the last three."""
    ni = a.size
    if ni < 3:
        raise ValueError("smooth: the input length must be longer than or equal to 3.")
    out = np.zeros(ni, dtype=np.complex128)
    out[0] = a[0]
    out[ni - 3:ni] = out[ni - 3:ni] + a[ni - 1] * np.array([0.2, 0.3, 0.5])
    for i in range(2, ni):               # reference i = 2 .. ni-1 (1-based): element a(i) = a[i-1]
        f = (1.0 + (ni - i) / (ni - 1.0)) * 0.5
        out[i - 2:i + 1] = out[i - 2:i + 1] + a[i - 1] * np.array([(1.0 - f) / 2.0, f, (1.0 - f) / 2.0])
    return out


def fftreat(s: Scalar, kit: Kit):
    """Smooth the far-field values of a scalar; ops:1002-1063.  The radial synthesis acts on the FFF array directly
    (the non-standard 'PFF' state of ops:1021), the tail rows ns.. of every retained (m,k) line are zeroed for
    m >= 1 beyond ns0, smoothed five times (over the whole tail INCLUDING the padding rows nr..nrdim-1), and the
    result goes back through the radial analysis."""
    ci = chop_index(s, kit)
    _require_fff(s, kit)
    p = kit.p
    nr = p.nr
    ns = nr * 3 // 4
    ns0 = min(ns + 4, nr)
    so = s.copy()
    ln = s.ln
    so.ln = 0.0
    delsqp(so, kit)
    rtrans_backward(so, kit)
    so.space = "PFF"
    first, second = k_ranges(ci, so.e.shape[2])
    kept = list(first) + [k for k in second if k not in set(first)]
    ncol = min(so.e.shape[1], ci.npc)
    fac = (1.0 - kit.x[:nr]) ** 2.0
    for mm in range(ncol):
        for kk in kept:
            so.e[:nr, mm, kk] = so.e[:nr, mm, kk] * fac
    so.e[ns0 - 1:, 1:ncol, :] = 0.0                       # loc_st(2) == 0: columns 2.. (m >= 1)
    for mm in range(ncol):
        for kk in kept:
            for _ in range(5):
                so.e[ns - 1:, mm, kk] = smooth(so.e[ns - 1:, mm, kk])
    for mm in range(ncol):
        for kk in kept:
            so.e[:nr, mm, kk] = so.e[:nr, mm, kk] / fac
    so.space = "FFF"
    rtrans_forward(so, kit)
    idelsqp(so, kit)
    so.ln = ln
    zeroat1(so, kit)
    s.e, s.ln, s.space = so.e, so.ln, so.space


# --------------------------------------------------------------------------- #
# time integrators
# --------------------------------------------------------------------------- #
def _hv_signed(p: Params) -> float:
    return p.hypervisc * (-1.0) ** (p.hyperpow // 2 + 1)


def _viscous_term(s: Scalar, kit: Kit) -> Scalar:
    """svis of fefe/abab/abcn; e.g. ops:1217-1230."""
    p = kit.p
    svis = s.copy()
    if p.hyperpow == 0:
        if p.visc < 5.0e-14:
            svis.e[...] = 0.0
        else:
            del2(svis, kit)
            svis.e = svis.e * p.visc
    else:
        helmp(svis, p.hyperpow, 0.0, p.visc / _hv_signed(p), kit)
        svis.e = svis.e * _hv_signed(p)
    return svis


def _check_fff(*fields):
    for f in fields:
        if f.space != "FFF":
            raise ValueError("fefe: all input scalars must be in FFF for time stepping")


def fefe(s: Scalar, nl: Scalar, dt: float, kit: Kit):
    """ops:1065-1094."""
    _check_fff(s, nl)
    svis = _viscous_term(s, kit)
    s.e = s.e + dt * (nl.e + svis.e)
    s.ln = s.ln + dt * (nl.ln + svis.ln)


def febe(s: Scalar, nl: Scalar, dt: float, kit: Kit):
    """forward Euler / backward Euler; ops:1157-1198."""
    _check_fff(s, nl)
    p = kit.p
    sh = s.copy()
    sh.e = s.e + dt * nl.e
    sh.ln = s.ln + dt * nl.ln
    if p.hyperpow == 0:
        if p.visc < 5.0e-14:
            raise ValueError("febe: inviscid case and no linear term in rhs. semi-implicit time adv is impossible")
        a = -1.0 / (dt * p.visc)
        sh.ln = a * sh.ln
        sh.e = a * sh.e
        ihelm(sh, a, kit)
    else:
        a = -1.0 / (dt * _hv_signed(p))
        b = p.visc / _hv_signed(p)
        sh.ln = a * sh.ln
        sh.e = a * sh.e
        ihelmp(sh, p.hyperpow, a, b, kit)
    s.e, s.ln = sh.e, sh.ln


def abcn(s: Scalar, s_p: Scalar, nl: Scalar, nl_p: Scalar, dt: float, kit: Kit):
    """Adams-Bashforth / Crank-Nicolson; ops:1200-1262.  Note ops:1256: s_p receives the NEW s."""
    _check_fff(s, nl, s_p, nl_p)
    p = kit.p
    sh = s.copy()
    sh.e = s.e + dt * (1.5 * nl.e - 0.5 * nl_p.e)
    sh.ln = s.ln + dt * (1.5 * nl.ln - 0.5 * nl_p.ln)
    svis = _viscous_term(s, kit)
    if p.hyperpow == 0:
        if p.visc < 5.0e-14:
            raise ValueError("abcn: inviscid case and no linear term in rhs. semi-implicit time adv is impossible")
        a = -2.0 / (dt * p.visc)
        sh.ln = a * (sh.ln + dt / 2.0 * svis.ln)
        sh.e = a * (sh.e + dt / 2.0 * svis.e)
        ihelm(sh, a, kit)
    else:
        a = -2.0 / (dt * _hv_signed(p))
        b = p.visc / _hv_signed(p)
        sh.ln = a * (sh.ln + dt / 2.0 * svis.ln)
        sh.e = a * (sh.e + dt / 2.0 * svis.e)
        ihelmp(sh, p.hyperpow, a, b, kit)
    s.e, s.ln = sh.e, sh.ln
    for dst, src in ((s_p, s), (nl_p, nl)):
        dst.e = src.e.copy(order="F")
        dst.ln, dst.space = src.ln, src.space
        dst.nrchop_offset, dst.npchop_offset, dst.nzchop_offset = (
            src.nrchop_offset, src.npchop_offset, src.nzchop_offset)


def abab(s: Scalar, s_p: Scalar, nl: Scalar, nl_p: Scalar, dt: float, kit: Kit, is_2nd_svis_p: bool = False):
    """Adams-Bashforth 2 on both terms; ops:1096-1155.  Quirk Q2 is reproduced: in the inviscid branch the reference
    zeroes `svis` a second time instead of `svis_p` (ops:1135), so svis_p stays a copy of s_p."""
    _check_fff(s, nl, s_p, nl_p)
    p = kit.p
    svis = _viscous_term(s, kit)
    if is_2nd_svis_p:
        svis_p = s_p.copy()
    elif p.hyperpow == 0 and p.visc < 5.0e-14:
        svis_p = s_p.copy()                      # ops:1135 writes svis%e, not svis_p%e
    else:
        svis_p = _viscous_term(s_p, kit)
    s.e = s.e + dt * (1.5 * (nl.e + svis.e) - 0.5 * (nl_p.e + svis_p.e))
    s.ln = s.ln + dt * (1.5 * (nl.ln + svis.ln) - 0.5 * (nl_p.ln + svis_p.ln))
    for dst, src in ((s_p, s), (nl_p, nl)):
        dst.e = src.e.copy(order="F")
        dst.ln, dst.space = src.ln, src.space
        dst.nrchop_offset, dst.npchop_offset, dst.nzchop_offset = (
            src.nrchop_offset, src.npchop_offset, src.nzchop_offset)


def helm(s: Scalar, alpha: float, kit: Kit):
    """(del^2 + alpha) s; ops:762-789.  The reference computes the result into a local `so` and never hands it back
    (no `s = so`, no function result), so its helm leaves s untouched apart from the transform to FFF.  This is the
    documented intent -- s <- del2(s) + alpha*s, ln <- alpha*ln (ops:785-786) -- i.e. the inverse of ihelm."""
    _require_fff(s, kit)
    so = s.copy()
    del2(so, kit)
    so.e = so.e + alpha * s.e
    so.ln = alpha * s.ln
    s.e, s.ln = so.e, so.ln


# --------------------------------------------------------------------------- #
# vector-field operations
# --------------------------------------------------------------------------- #
def vecprod(vr, vp, vz, ur, up, uz, kit: Kit):
    """pointwise v x u on the Re and Im lanes separately; ops:1264-1306."""
    for f in (vr, vp, vz, ur, up, uz):
        if f.space != "PPP":
            raise ValueError("vector_product: to compute v x u, all components must be in PPP")
    p = kit.p
    sl = (slice(0, p.nr), slice(0, p.np // 2), slice(0, p.nz))
    a1, a2, a3 = vr.e[sl].real, vp.e[sl].real, vz.e[sl].real
    b1, b2, b3 = ur.e[sl].real, up.e[sl].real, uz.e[sl].real
    c1, c2, c3 = vr.e[sl].imag, vp.e[sl].imag, vz.e[sl].imag
    d1, d2, d3 = ur.e[sl].imag, up.e[sl].imag, uz.e[sl].imag
    outr = np.zeros_like(vr.e)
    outp = np.zeros_like(vr.e)
    outz = np.zeros_like(vr.e)
    outr[sl] = (a2 * b3 - a3 * b2) + 1j * (c2 * d3 - c3 * d2)
    outp[sl] = (a3 * b1 - a1 * b3) + 1j * (c3 * d1 - c1 * d3)
    outz[sl] = (a1 * b2 - a2 * b1) + 1j * (c1 * d2 - c2 * d1)
    vr.e, vp.e, vz.e = outr, outp, outz


def _eo_fold(b: np.ndarray, nr: int):
    njh = nr // 2
    top = b[:njh]
    bot = b[nr - 1: njh - 1: -1]
    return top + bot, top - bot


def eomul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """ops:2067-2104: odd rows (1-based) against the even fold, even rows against the odd fold."""
    be, bo = _eo_fold(b, b.shape[0])
    c = np.zeros((a.shape[0],) + b.shape[1:], dtype=np.complex128)
    c[0::2] = a[0::2, :] @ be
    if a.shape[0] > 1:
        c[1::2] = a[1::2, :] @ bo
    return c


def oemul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """ops:2108-2145."""
    be, bo = _eo_fold(b, b.shape[0])
    c = np.zeros((a.shape[0],) + b.shape[1:], dtype=np.complex128)
    c[0::2] = a[0::2, :] @ bo
    if a.shape[0] > 1:
        c[1::2] = a[1::2, :] @ be
    return c


def vec2tp(vr: Scalar, vp: Scalar, vz: Scalar, psi: Scalar, chi: Scalar, kit: Kit):
    """Toroidal-poloidal projection; ops:1308-1453."""
    for f in (vr, vp, vz):
        if f.space != "PPP":
            raise ValueError("vector_projection: vr must be in PPP")
    if psi.space != "FFF":
        raise ValueError("vector_projection: psi must be in FFF")
    if chi.space != "FFF":
        raise ValueError("vector_projection: chi must be in FFF")
    p = kit.p
    nr, nrh = p.nr, p.nr // 2
    r = p.ell * np.sqrt((1.0 + kit.x) / (1.0 - kit.x))
    psi.e[...] = 0.0
    chi.e[...] = 0.0
    psi.ln = 0.0
    chi.ln = 0.0
    ur, up, uz = vr.copy(), vp.copy(), vz.copy()
    ur.e[:nr] = ur.e[:nr] * r[:, None, None]
    up.e[:nr] = up.e[:nr] * r[:, None, None]

    trans(ur, "PFP", kit)
    if p.nz > 1:
        vertical_fft_forward(ur, kit)
    ur.space = "PFF"

    trans(up, "FFF", kit)
    inf = calcat1(up, kit)
    psi.ln = -1.0 / 2.0 * float(inf[0].real)
    rtrans_backward(up, kit)
    up.space = "PFF"
    up.e[:nr, 0, 0] = up.e[:nr, 0, 0] + psi.ln * (1.0 + kit.x)

    trans(uz, "PFP", kit)
    if p.nz > 1:
        vertical_fft_forward(uz, kit)
    uz.space = "PFF"

    ur.chop_offset(3)
    up.chop_offset(3)
    uz.chop_offset(3)
    ci = chop_index(ur, kit)
    first, second = k_ranges(ci, psi.e.shape[2])
    fff = (1.0 - kit.x[:nrh] ** 2.0) / kit.w[:nrh]

    for mm in range(min(ur.e.shape[1], ci.npc)):
        nn = int(ci.nrcs[mm])
        mv = int(kit.m[mm])
        xb = leg_xxdx(mv, nn + 1, kit)
        pfm = kit.pf[:nrh, : nn + 1, mm]
        # pfd = pf .mul. xxdx (dense x band, bops:135-160): pfd(:,j) = sum_k pf(:,k) X(k,j), k ascending
        pfd = np.zeros((nrh, nn + 1))
        pfd[:, 1:] = pfd[:, 1:] + pfm[:, :-1] * xb[1][:-1][None, :]      # X(j-1, j)
        pfd[:, :-1] = pfd[:, :-1] + pfm[:, 1:] * xb[-1][1:][None, :]     # X(j+1, j)
        v = np.zeros((nn, nrh))
        d = np.zeros((nn, nrh))
        t = np.zeros((nn, nrh))
        for i in range(nn):
            n = max(1, mv + i)
            v[i, :] = pfm[:, i] / n / (n + 1) / fff
            d[i, :] = pfd[:, i] / n / (n + 1) / fff
            t[i, :] = pfm[:, i] * kit.w[:nrh]
        for rng in (first, second):
            if not rng:
                continue
            ks = slice(rng[0], rng[-1] + 1)
            kv = kit.ak[ks][None, :]
            ure, upe, uze = ur.e[:nr, mm, ks], up.e[:nr, mm, ks], uz.e[:nr, mm, ks]
            w1 = eomul(v, ure) if mv != 0 else np.zeros((nn, ure.shape[1]), complex)
            w2 = oemul(d, upe)
            psi.e[:nn, mm, ks] = -1j * mv * w1 - w2
            w1 = oemul(d, ure)
            w2 = eomul(v, upe) if mv != 0 else np.zeros((nn, ure.shape[1]), complex)
            ct = eomul(t, uze)
            chi.e[:nn, mm, ks] = 1j * kv * w1 + mv * kv * w2 - ct

    idel2_proln(chi, kit)
    psi.chop_offset(0)
    chi.chop_offset(0)
    chop(psi, kit)
    chop(chi, kit)
    zeroat1(psi, kit)
    zeroat1(chi, kit)


def tp2vec(psi: Scalar, chi: Scalar, kit: Kit):
    """Solenoidal field from its toroidal/poloidal scalars; ops:1455-1545.  Returns (vr, vp, vz) in PPP."""
    if psi.space != "FFF":
        raise ValueError("vector_projection: psi must be in FFF")
    if chi.space != "FFF":
        raise ValueError("vector_projection: chi must be in FFF")
    p = kit.p
    nr = p.nr
    r = p.ell * np.sqrt((1.0 + kit.x) / (1.0 - kit.x))
    ur, up, uz = chi.copy(), psi.copy(), chi.copy()
    for f in (ur, up, uz):
        f.chop_offset(3)
    ci = chop_index(ur, kit)
    xxdx(ur, kit)
    xxdx(up, kit)
    nzdim = ur.e.shape[2]
    kmask = np.array([(kk + 1 <= ci.nzc) or (kk + 1 >= ci.nzcu) for kk in range(nzdim)])
    kv = kit.ak[None, :nzdim]
    for mm in range(min(ur.e.shape[1], ci.npc)):
        nn = int(ci.nrcs[mm])
        mv = int(kit.m[mm])
        new_ur = 1j * mv * psi.e[:nn, mm, :] + 1j * kv * ur.e[:nn, mm, :]
        new_up = -up.e[:nn, mm, :] - mv * kv * uz.e[:nn, mm, :]
        ur.e[:nn, mm, kmask] = new_ur[:, kmask]
        up.e[:nn, mm, kmask] = new_up[:, kmask]
    ur.chop_offset(0)
    up.chop_offset(0)
    chop(ur, kit)
    chop(up, kit)
    trans(ur, "PPP", kit)
    trans(up, "PPP", kit)
    ur.e[:nr] = ur.e[:nr] / r[:, None, None]
    up.e[:nr] = up.e[:nr] / r[:, None, None]
    del2h(uz, kit)
    uz.e = -uz.e
    uz.chop_offset(0)
    chop(uz, kit)
    trans(uz, "PPP", kit)
    return ur, up, uz


def tp2curlvec(psi: Scalar, chi: Scalar, kit: Kit):
    """curl of the field: tp2vec(-del2(chi), psi); ops:1547-1560."""
    mdel2chi = chi.copy()
    mdel2chi.space = "FFF"
    del2(mdel2chi, kit)
    mdel2chi.e = -mdel2chi.e
    return tp2vec(mdel2chi, psi, kit)


# --------------------------------------------------------------------------- #
# the 3-D vortex application (apps/vortical_flow_3d.f90)
# --------------------------------------------------------------------------- #
def qvort_dist_tp(kit: Kit, q: float = 1.0, ran_noise: float = 0.0):
    """Two q-vortices at x = -2, +2; apps/vortical_flow_3d.f90:258-326 with ran_noise = 0
    (gfortran's random_number stream is not reproducible elsewhere)."""
    if ran_noise != 0.0:
        raise ValueError("qvort_dist_tp: only ran_noise = 0 is reproducible")
    p = kit.p
    nr, nph, nz = p.nr, p.np // 2, p.nz
    pang = np.array([2.0 * PI / p.np * i for i in range(p.np + 1)])
    out = []
    for amp in (2.0, 1.0 / q):
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
            if amp == 2.0:
                acc = acc + (-np.exp(-(rr ** 2.0)) * 2.0 / den + 1j * (-np.exp(-(ri ** 2.0)) * 2.0 / den))
            else:
                acc = acc + (-np.exp(-(rr ** 2.0)) / q / den + 1j * (-np.exp(-(ri ** 2.0)) / q / den))
        glb[:nr, :nph, :nz] = acc[:, :, None]
        s = Scalar(e=glb, space="PPP")
        trans(s, "FFF", kit)
        idelsqp(s, kit)
        zeroat1(s, kit)
        out.append(s)
    return out[0], out[1]


def uniform_z_fld(kit: Kit, b: float = -0.5) -> Scalar:
    """apps/vortical_flow_3d.f90:328-351."""
    p = kit.p
    s = scalar_init(kit, "PPP")
    s.e[: p.nr, : p.np // 2, : p.nz] = complex(b, b)
    return s


def advection_rhs(psi: Scalar, chi: Scalar, uz: Scalar, kit: Kit):
    """apps/vortical_flow_3d.f90:353-395.  Dealiases psi, chi in place; returns (nlpsi, nlchi)."""
    dealias(psi, kit)
    dealias(chi, kit)
    vr, vp, vz = tp2vec(psi, chi, kit)
    vz.e = vz.e + uz.e
    wr, wp, wz = tp2curlvec(psi, chi, kit)
    vecprod(vr, vp, vz, wr, wp, wz, kit)
    nlpsi = scalar_init(kit, "FFF")
    nlchi = scalar_init(kit, "FFF")
    vec2tp(vr, vp, vz, nlpsi, nlchi, kit)
    nlpsi.ln = 0.0
    nlchi.ln = 0.0
    dealias(nlpsi, kit)
    dealias(nlchi, kit)
    return nlpsi, nlchi


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
    gain_psi: float = 0.0
    gain_chi: float = 0.0


def vortex_bootstrap(kit: Kit, dt: float, psi: Scalar, chi: Scalar, uz: Scalar) -> VortexState:
    """Richardson-extrapolated FEBE first step; apps/vortical_flow_3d.f90:116-147."""
    nlpsi, nlchi = advection_rhs(psi, chi, uz, kit)
    psi_prev, nlpsi_prev = psi.copy(), nlpsi.copy()
    chi_prev, nlchi_prev = chi.copy(), nlchi.copy()
    psi_rich, chi_rich = psi.copy(), chi.copy()
    nlpsi, nlchi = advection_rhs(psi_rich, chi_rich, uz, kit)
    febe(psi_rich, nlpsi, dt, kit)
    febe(chi_rich, nlchi, dt, kit)
    for _ in range(2):
        nlpsi, nlchi = advection_rhs(psi, chi, uz, kit)
        febe(psi, nlpsi, dt / 2.0, kit)
        febe(chi, nlchi, dt / 2.0, kit)
    psi.e = 2.0 / 1.0 * psi.e - 1.0 / 1.0 * psi_rich.e
    chi.e = 2.0 / 1.0 * chi.e - 1.0 / 1.0 * chi_rich.e
    st = VortexState(psi, chi, nlpsi, nlchi, psi_prev, chi_prev, nlpsi_prev, nlchi_prev, uz)
    dealias(psi, kit)
    dealias(chi, kit)
    st.gain_psi = svv_filter(psi, kit, st.gain_psi)
    st.gain_chi = svv_filter(chi, kit, st.gain_chi)
    zeroat1(psi, kit)
    zeroat1(chi, kit)
    st.nlpsi, st.nlchi = advection_rhs(psi, chi, uz, kit)
    return st


def vortex_step(st: VortexState, kit: Kit, dt: float):
    """One ABCN step of the main loop; apps/vortical_flow_3d.f90:160-180."""
    abcn(st.psi, st.psi_prev, st.nlpsi, st.nlpsi_prev, dt, kit)
    abcn(st.chi, st.chi_prev, st.nlchi, st.nlchi_prev, dt, kit)
    dealias(st.psi, kit)
    dealias(st.chi, kit)
    st.gain_psi = svv_filter(st.psi, kit, st.gain_psi)
    st.gain_chi = svv_filter(st.chi, kit, st.gain_chi)
    zeroat1(st.psi, kit)
    zeroat1(st.chi, kit)
    st.nlpsi, st.nlchi = advection_rhs(st.psi, st.chi, st.uz, kit)
    if not (np.all(np.isfinite(st.psi.e)) and np.all(np.isfinite(st.chi.e))):
        raise FloatingPointError("ERROR: non-finite vortex state")   # check_stability, :397-409


# --------------------------------------------------------------------------- #
# pencil decomposition / exchange (index bookkeeping only; used by the gloo tests)
# --------------------------------------------------------------------------- #
def decompose(nsize: int, nprocs: int, proc_num: int):
    """submodules/mlegs_envir_mpi.f90:6-31 -> (local size, 0-based start)."""
    q, r = divmod(nsize, nprocs)
    if r > proc_num:
        return q + 1, (q + 1) * proc_num
    return q, q * proc_num + r


def exchange_global(blocks_old, glb_sz, axis_old: int, axis_new: int):
    """Reference meaning of scalar_exchange on P ranks of one communicator
    (submodules/mlegs_scalar_dist.f90:6-67): before, every rank holds all of axis_old and its
    ``decompose`` share of axis_new; after, all of axis_new and a share of axis_old.
    ``blocks_old``: list over ranks of local arrays.  Axes are 1-based like the reference."""
    P = len(blocks_old)
    full = np.concatenate(blocks_old, axis=axis_new - 1)
    assert full.shape == tuple(glb_sz)
    out = []
    for rk in range(P):
        sz, st = decompose(glb_sz[axis_old - 1], P, rk)
        sl = [slice(None)] * 3
        sl[axis_old - 1] = slice(st, st + sz)
        out.append(np.asfortranarray(full[tuple(sl)]))
    return out
