"""Third-party pins of the oracle's tables (the reference holds no numbers, see test_oracle_analytic.py): the quadrature
against numpy's Gauss-Legendre rule and the orthonormal mapped associated Legendre functions against mpmath's
`legenp` (type 2: Ferrers functions with the Condon-Shortley phase) over a spread of degrees, orders and nodes --
not only the two closed forms of the tutorial."""
import math

import mpmath
import numpy as np
import pytest

from oracle import mlegs_oracle as mo


@pytest.fixture(scope="module")
def kit3d():
    p = mo.Params(nr=48, np=32, nz=4, nrchop=44, npchop=17, nzchop=3, ell=3.0, zlen=2.0 * math.pi,
                  visc=1.0e-3, hyperpow=0, hypervisc=0.0)
    return mo.kit_init(p)


def test_quadrature_is_gauss_legendre(kit3d):
    x, w = np.polynomial.legendre.leggauss(kit3d.p.nr)
    assert np.max(np.abs(np.sort(kit3d.x) - x)) < 4e-16
    order = np.argsort(kit3d.x)
    # the weights come out of the reference's double-precision Newton iteration, w = 2 / ((1 - z^2) P_n'(z)^2)
    # (sinit:181-219): a few 1e-15 absolute, 1e-12 relative next to the end points where the weights are small
    assert np.max(np.abs(kit3d.w[order] - w)) < 2e-14
    assert np.max(np.abs(kit3d.w[order] - w) / w) < 5e-12
    # r = ell sqrt((1 + x) / (1 - x)): the map of docs/tutorial/transformation.md
    r = kit3d.p.ell * np.sqrt((1.0 + kit3d.x) / (1.0 - kit3d.x))
    assert np.max(np.abs(kit3d.r - r) / r) < 1e-14


def test_basis_functions_against_mpmath_legenp(kit3d):
    mpmath.mp.dps = 40
    nrh = kit3d.p.nr // 2
    rng = np.random.default_rng(3)
    worst = 0.0
    checked = 0
    for m in (0, 1, 2, 5, 9, 16):
        nn = int(kit3d.chops[m])
        for j in sorted(set([0, 1, 2, nn // 2, nn - 2, nn - 1]) & set(range(nn))):
            n = m + j
            norm = mpmath.sqrt(mpmath.mpf(2 * n + 1) / 2 * mpmath.factorial(n - m) / mpmath.factorial(n + m))
            for i in rng.choice(nrh, size=4, replace=False):
                ref = norm * mpmath.legenp(n, m, mpmath.mpf(float(kit3d.x[i])), type=2)
                got = kit3d.pf[i, j, m]
                scale = max(abs(float(ref)), 1e-3)
                worst = max(worst, abs(got - float(ref)) / scale)
                checked += 1
    assert checked > 100
    # the nodes themselves are double-precision numbers (sinit:187): evaluating the exact function AT the stored node
    # leaves rounding of the 50-digit recurrence's result only
    assert worst < 5e-13, worst


def _radial_values(kit, s):
    """Radial synthesis of every (m, k) column of an FFF coefficient array (the third axis is only carried along)."""
    s.space = "FFP"
    mo.rtrans_backward(s, kit)
    return s.e


@pytest.mark.parametrize("m", [0, 1, 3, 8])
def test_band_operators_against_legendre_identities(kit3d, m):
    """xxdx = r d/dr and del2 = (1/r) d/dr (r d/dr) - m^2/r^2 - ak(k)^2 of a single basis function: the oracle's band
    tables (the coefficient formulas of sdiff:6-152, which the device kernels tabulate too) followed by its radial
    synthesis, against derivatives built from mpmath's legenp with the textbook identities only --
    (1 - x^2) P' = (n + 1) x P_n^m - (n - m + 1) P_{n+1}^m (DLMF 14.10.5), P'' from Legendre's equation, and the chain
    rule through x = (r^2 - L^2) / (r^2 + L^2)."""
    mpmath.mp.dps = 40
    kit = kit3d
    ell = mpmath.mpf(kit.p.ell)
    nn = int(kit.chops[m])
    nodes = [0, 3, 9, 17, 24, 31, 40, 47]
    for j in (0, 1, 6, nn - 4):
        n = m + j
        norm = mpmath.sqrt(mpmath.mpf(2 * n + 1) / 2 * mpmath.factorial(n - m) / mpmath.factorial(n + m))
        fv, d1v, d2v, rv = [], [], [], []
        for i in nodes:
            x = mpmath.mpf(float(kit.x[i]))
            r = ell * mpmath.sqrt((1 + x) / (1 - x))
            p0 = mpmath.legenp(n, m, x, type=2)
            p1 = mpmath.legenp(n + 1, m, x, type=2)
            dp = ((n + 1) * x * p0 - (n - m + 1) * p1) / (1 - x * x)
            ddp = (2 * x * dp - (n * (n + 1) - mpmath.mpf(m * m) / (1 - x * x)) * p0) / (1 - x * x)
            den = r * r + ell * ell
            xr = 4 * r * ell * ell / den ** 2
            xrr = 4 * ell * ell * (ell * ell - 3 * r * r) / den ** 3
            rv.append(r)
            fv.append(norm * p0)
            d1v.append(norm * dp * xr)
            d2v.append(norm * (ddp * xr * xr + dp * xrr))
        for k in (0, 1):
            ak = mpmath.mpf(float(kit.ak[k]))
            for op in ("xxdx", "del2"):
                s = mo.scalar_init(kit, "FFF")
                s.e[j, m, k] = 1.0
                getattr(mo, op)(s, kit)
                got = _radial_values(kit, s)[:, m, k]
                if op == "xxdx":
                    exact = np.array([float(r * d1) for r, d1 in zip(rv, d1v)])
                else:
                    exact = np.array([float(d2 + d1 / r - (mpmath.mpf(m * m) / (r * r) + ak * ak) * f)
                                      for r, f, d1, d2 in zip(rv, fv, d1v, d2v)])
                scale = max(np.max(np.abs(exact)), 1e-30)
                err = np.max(np.abs(got[nodes].real - exact)) / scale
                assert err < 2e-11, (op, m, j, k, err)
                assert np.max(np.abs(got[nodes].imag)) <= 1e-13 * scale


# ---- the whole transform from its published definition -------------------------------------------------------------
# docs/tutorial/initialization.md:154 of the reference states what the coefficients mean:
#     s(r, phi, z) = sum_kappa sum_m sum_{n >= |m|} s_n^{m kappa} P_{L_n}^m(r) exp(i m phi + i kappa z).
# The two tests below evaluate that triple sum (and its inverse, the quadrature) term by term with mpmath's legenp at
# EVERY radial node, numpy's exp and numpy's Gauss-Legendre weights -- no FFT, no even/odd fold, none of the oracle's
# tables -- and compare with the oracle's trans(): this pins the exponent signs, where the 1/N factors sit, the packing
# of phi pairs into one complex number, the FFT order of k and the parity fold of the radial stage.

@pytest.fixture(scope="module")
def kit_small():
    p = mo.Params(nr=16, np=8, nz=4, nrchop=16, npchop=5, nzchop=3, ell=2.5, zlen=3.0,
                  visc=1.0e-3, hyperpow=0, hypervisc=0.0)
    return mo.kit_init(p)


@pytest.fixture(scope="module")
def basis_small(kit_small):
    return _basis_at_all_nodes(kit_small)


def _basis_at_all_nodes(kit):
    """B[i, j, m] = orthonormal P_{m+j}^m(x_i) at all nr nodes, from mpmath (30 digits)."""
    mpmath.mp.dps = 30
    nr, npc = kit.p.nr, kit.chopp
    B = np.zeros((nr, kit.p.nrchop, npc))
    for m in range(npc):
        for j in range(int(kit.chops[m])):
            n = m + j
            norm = mpmath.sqrt(mpmath.mpf(2 * n + 1) / 2 * mpmath.factorial(n - m) / mpmath.factorial(n + m))
            for i in range(nr):
                B[i, j, m] = float(norm * mpmath.legenp(n, m, mpmath.mpf(float(kit.x[i])), type=2))
    return B


def _signed_k(nz):
    k = np.arange(nz)
    return np.where(k <= nz // 2, k, k - nz)     # FFT order: 0, 1, ..., nz/2, -(nz/2-1), ..., -1


def test_synthesis_equals_the_published_triple_sum(kit_small, basis_small):
    kit = kit_small
    nr, npts, nz = kit.p.nr, kit.p.np, kit.p.nz
    nph = npts // 2
    rng = np.random.default_rng(11)
    a = np.zeros(kit.glb_sz, dtype=complex)
    for m in range(kit.chopp):
        nn = int(kit.chops[m])
        a[:nn, m, :nz] = rng.standard_normal((nn, nz)) + 1j * rng.standard_normal((nn, nz))
    s = mo.Scalar(e=np.asfortranarray(a.copy()), space="FFF")
    mo.trans(s, "PPP", kit)
    got = s.e[:nr, :nph, :nz]
    phys_got = np.empty((nr, npts, nz))
    phys_got[:, 0::2, :] = got.real          # Re = phi_{2q}, Im = phi_{2q+1}
    phys_got[:, 1::2, :] = got.imag

    B = basis_small
    kk = _signed_k(nz)
    ez = np.exp(2j * np.pi * np.outer(kk, np.arange(nz)) / nz)          # [k, l]  exp(i kappa z_l), z_l = zlen l / nz
    ephi = np.exp(2j * np.pi * np.outer(np.arange(kit.chopp), np.arange(npts)) / npts)   # [m, p]  exp(i m phi_p)
    # c[i, m, l] = sum_k sum_j a[j, m, k] B[i, j, m] exp(i kappa z_l)
    c = np.einsum("ijm,jmk,kl->iml", B, a[: kit.p.nrchop, : kit.chopp, :nz], ez)
    # the field is real: the m < 0 terms are the conjugates of the m > 0 ones; m = 0 and the Nyquist m = np/2 appear once
    want = np.zeros((nr, npts, nz))
    for m in range(kit.chopp):
        term = (c[:, m, None, :] * ephi[m][None, :, None]).real
        want += term if m in (0, nph) else 2.0 * term
    assert np.max(np.abs(phys_got - want)) / np.max(np.abs(want)) < 1e-13


def test_analysis_equals_the_quadrature_of_the_definition(kit_small, basis_small):
    kit = kit_small
    nr, npts, nz = kit.p.nr, kit.p.np, kit.p.nz
    nph = npts // 2
    rng = np.random.default_rng(12)
    phys = rng.standard_normal((nr, npts, nz))
    e = np.zeros(kit.glb_sz, dtype=complex)
    e[:nr, :nph, :nz] = phys[:, 0::2, :] + 1j * phys[:, 1::2, :]
    s = mo.Scalar(e=np.asfortranarray(e), space="PPP")
    mo.trans(s, "FFF", kit)

    xg, wg = np.polynomial.legendre.leggauss(nr)          # third-party rule, matched to the kit's node order
    order = np.argsort(kit.x)
    w = np.empty(nr)
    w[order] = wg
    assert np.max(np.abs(kit.x[order] - xg)) < 4e-16
    B = basis_small
    kk = _signed_k(nz)
    ez = np.exp(-2j * np.pi * np.outer(np.arange(nz), kk) / nz) / nz                       # [l, k]
    ephi = np.exp(-2j * np.pi * np.outer(np.arange(npts), np.arange(kit.chopp)) / npts) / npts   # [p, m]
    want = np.einsum("i,ijm,ipl,pm,lk->jmk", w, B, phys, ephi, ez)
    got = s.e[: kit.p.nrchop, : kit.chopp, :nz]
    for m in range(kit.chopp):
        nn = int(kit.chops[m])
        assert np.max(np.abs(got[:nn, m] - want[:nn, m])) / np.max(np.abs(want[:nn, m])) < 1e-13, m
        assert not np.any(got[nn:, m])                       # beyond the triangular truncation: exact zeros
