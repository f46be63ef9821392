"""The spectral transform evaluated from its PUBLISHED definition, term by term, with third-party functions only.

docs/tutorial/initialization.md:154 of the reference states what the coefficients of a scalar mean:

    s(r, phi, z) = sum_kappa sum_m sum_{n >= |m|} s_n^{m kappa} P_{L_n}^m(r) exp(i m phi + i kappa z),

P_{L_n}^m the orthonormal associated Legendre function of the mapped coordinate x = (r^2 - L^2) / (r^2 + L^2).  The
functions below evaluate that triple sum and its inverse (Gauss-Legendre quadrature in x, plain sums in phi and z) with
mpmath's / scipy's Legendre functions, numpy's exp and numpy's Gauss-Legendre weights: no FFT, no even/odd fold, none of
the oracle's or the device library's tables.  Shared by tests/test_oracle_independent.py (the oracle against the
definition, CPU) and tests/test_zz_gpu_from_definition.py (the CUDA path against the definition, without the oracle in
between)."""
from __future__ import annotations

import math

import numpy as np


def basis_at_all_nodes(x, nrchop: int, npc: int, backend: str = "mpmath") -> np.ndarray:
    """B[i, j, m] = sqrt((2n+1)/2 (n-m)!/(n+m)!) P_n^m(x_i), n = m + j < nrchop (Ferrers function with the Condon-Shortley
    phase), at EVERY node x_i."""
    x = np.asarray(x, dtype=np.float64)
    B = np.zeros((x.size, nrchop, npc))
    if backend == "mpmath":
        import mpmath
        mpmath.mp.dps = 30
        for m in range(npc):
            for j in range(max(nrchop - m, 0)):
                n = m + j
                norm = mpmath.sqrt(mpmath.mpf(2 * n + 1) / 2 * mpmath.factorial(n - m) / mpmath.factorial(n + m))
                for i in range(x.size):
                    B[i, j, m] = float(norm * mpmath.legenp(n, m, mpmath.mpf(float(x[i])), type=2))
    elif backend == "scipy":
        from scipy.special import lpmv
        for m in range(npc):
            for j in range(max(nrchop - m, 0)):
                n = m + j
                norm = math.sqrt((2 * n + 1) / 2.0 * math.factorial(n - m) / math.factorial(n + m))
                B[:, j, m] = norm * lpmv(m, n, x)
    else:
        raise ValueError(backend)
    return B


def signed_k(nz: int) -> np.ndarray:
    """axial wavenumber index of every plane in FFT order: 0, 1, ..., nz/2, -(nz/2 - 1), ..., -1"""
    k = np.arange(nz)
    return np.where(k <= nz // 2, k, k - nz)


def unpack_ppp(e: np.ndarray, nr: int, npts: int, nz: int) -> np.ndarray:
    """PPP memory image -> real field f[i, p, l]: one complex number holds phi_{2q} (Re) and phi_{2q+1} (Im)."""
    nph = npts // 2
    f = np.empty((nr, npts, nz))
    f[:, 0::2, :] = e[:nr, :nph, :nz].real
    f[:, 1::2, :] = e[:nr, :nph, :nz].imag
    return f


def pack_ppp(f: np.ndarray, glb_sz) -> np.ndarray:
    nr, npts, nz = f.shape
    e = np.zeros(glb_sz, dtype=np.complex128, order="F")
    e[:nr, : npts // 2, :nz] = f[:, 0::2, :] + 1j * f[:, 1::2, :]
    return e


def synthesis_by_definition(a: np.ndarray, B: np.ndarray, npts: int, nz: int) -> np.ndarray:
    """The real field of the coefficients a[j, m, k] (m = 0 .. npc-1 <= np/2, k in FFT order) on the grid
    phi_p = 2 pi p / np, z_l = zlen l / nz.  The m < 0 terms of a real field are the conjugates of the m > 0 ones; m = 0
    and the azimuthal Nyquist m = np/2 appear once."""
    nr, nrchop, npc = B.shape
    nph = npts // 2
    ez = np.exp(2j * np.pi * np.outer(signed_k(nz), np.arange(nz)) / nz)             # [k, l]  exp(i kappa z_l)
    ephi = np.exp(2j * np.pi * np.outer(np.arange(npc), np.arange(npts)) / npts)     # [m, p]  exp(i m phi_p)
    c = np.einsum("ijm,jmk,kl->iml", B, a[:nrchop, :npc, :nz], ez)                   # radial + axial sums
    f = np.zeros((nr, npts, nz))
    for m in range(npc):
        term = (c[:, m, None, :] * ephi[m][None, :, None]).real
        f += term if m in (0, nph) else 2.0 * term
    return f


def analysis_by_definition(f: np.ndarray, B: np.ndarray, x_nodes: np.ndarray) -> np.ndarray:
    """a[j, m, k] = sum_i w_i B[i, j, m] (1/np) sum_p (1/nz) sum_l f[i, p, l] exp(-i m phi_p - i kappa z_l), the weights
    from numpy's Gauss-Legendre rule matched to the node order of x_nodes."""
    nr, npts, nz = f.shape
    npc = B.shape[2]
    xg, wg = np.polynomial.legendre.leggauss(nr)
    order = np.argsort(x_nodes)
    assert np.max(np.abs(np.asarray(x_nodes)[order] - xg)) < 4e-16
    w = np.empty(nr)
    w[order] = wg
    ez = np.exp(-2j * np.pi * np.outer(np.arange(nz), signed_k(nz)) / nz) / nz                  # [l, k]
    ephi = np.exp(-2j * np.pi * np.outer(np.arange(npts), np.arange(npc)) / npts) / npts        # [p, m]
    return np.einsum("i,ijm,ipl,pm,lk->jmk", w, B, f, ephi, ez, optimize=True)


def random_triangular(glb_sz, nrchop: int, npc: int, nz: int, seed: int) -> np.ndarray:
    """random complex coefficients inside the triangular truncation n < nrchop, zeros elsewhere (FFF memory image)"""
    rng = np.random.default_rng(seed)
    a = np.zeros(glb_sz, dtype=np.complex128, order="F")
    for m in range(npc):
        nn = max(nrchop - m, 0)
        a[:nn, m, :nz] = rng.standard_normal((nn, nz)) + 1j * rng.standard_normal((nn, nz))
    return a
