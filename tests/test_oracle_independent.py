"""Pins of the oracle that do not come from the oracle (the reference holds no numbers, see test_oracle_analytic.py).

Third-party evaluations: the quadrature against numpy's Gauss-Legendre rule; the orthonormal mapped associated Legendre
functions against mpmath's `legenp` (type 2: Ferrers functions with the Condon-Shortley phase) over a spread of degrees,
orders and nodes; the xxdx / del2 band tables against derivatives built from `legenp` and textbook identities.

The reference's own PUBLISHED definitions, evaluated term by term: the transform both ways against the triple sum of
docs/tutorial/initialization.md:154; tp2vec against the curl formulas of docs/tutorial/vector_field.md:18,58-62 (and
vec2tp as its inverse on a field with azimuthal and axial structure); ihelm / ihelmp as inverses of the operator applied
as repeated del2; febe / abcn against the discretisation of docs/tutorial/time_integration.md:160-169, 381-390."""
import math

import mpmath
import numpy as np
import pytest

from oracle import mlegs_oracle as mo
import definition as dfn


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
def kit_mid():
    # the lengths the register-resident FFT kernels of the device library start at (np/2 = nz = 32)
    p = mo.Params(nr=32, np=64, nz=32, nrchop=32, npchop=5, nzchop=17, ell=4.0, zlen=2.0 * math.pi,
                  visc=1.0e-3, hyperpow=0, hypervisc=0.0)
    return mo.kit_init(p)


def _from_definition_both_ways(kit, backend, tol):
    nr, npts, nz = kit.p.nr, kit.p.np, kit.p.nz
    B = dfn.basis_at_all_nodes(kit.x, kit.p.nrchop, kit.chopp, backend)
    # synthesis
    a = dfn.random_triangular(kit.glb_sz, kit.p.nrchop, kit.chopp, nz, seed=11)
    s = mo.Scalar(e=a.copy(order="F"), space="FFF")
    mo.trans(s, "PPP", kit)
    want = dfn.synthesis_by_definition(a, B, npts, nz)
    assert np.max(np.abs(dfn.unpack_ppp(s.e, nr, npts, nz) - want)) / np.max(np.abs(want)) < tol
    # analysis
    f = np.random.default_rng(12).standard_normal((nr, npts, nz))
    s = mo.Scalar(e=dfn.pack_ppp(f, kit.glb_sz), space="PPP")
    mo.trans(s, "FFF", kit)
    want = dfn.analysis_by_definition(f, B, kit.x)
    got = s.e[: kit.p.nrchop, : kit.chopp, :nz]
    for m in range(kit.chopp):
        nn = int(kit.chops[m])
        assert np.max(np.abs(got[:nn, m] - want[:nn, m])) / np.max(np.abs(want[:nn, m])) < tol, m
        assert not np.any(got[nn:, m])                       # beyond the triangular truncation: exact zeros


def test_transform_equals_the_published_triple_sum_mpmath(kit_small):
    _from_definition_both_ways(kit_small, "mpmath", 1e-13)


def test_transform_equals_the_published_triple_sum_scipy(kit_small, kit_mid):
    # scipy's lpmv as the third-party Legendre function: fast enough for the size the GPU test of the same name runs
    _from_definition_both_ways(kit_small, "scipy", 1e-12)
    _from_definition_both_ways(kit_mid, "scipy", 1e-12)


# ---- the toroidal-poloidal reconstruction from its published definition ---------------------------------------------
# docs/tutorial/vector_field.md:18,58-62:  V = curl(psi e_z) + curl curl(chi e_z), i.e.
#     V_r = (1/r) d psi/d phi + d^2 chi/(dr dz),  V_phi = - d psi/dr + (1/r) d^2 chi/(d phi dz),  V_z = - del_perp^2 chi.
# One mode each for psi and chi, differentiated by hand (mpmath legenp + DLMF 14.10.5 + Legendre's equation + the chain
# rule through the map), against the oracle's tp2vec (band tables, chop offsets, three backward transforms, the 1/r).

def _mode_and_derivatives(kit, n, m, nodes):
    """orthonormal P_{L_n}^m and its first two r-derivatives at the given node indices (mpmath numbers)."""
    ell = mpmath.mpf(kit.p.ell)
    norm = mpmath.sqrt(mpmath.mpf(2 * n + 1) / 2 * mpmath.factorial(n - m) / mpmath.factorial(n + m))
    out = []
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
        out.append((r, norm * p0, norm * dp * xr, norm * (ddp * xr * xr + dp * xrr)))
    return out


@pytest.mark.parametrize("m,k,jpsi,jchi", [(3, 1, 2, 1), (1, 3, 0, 4), (5, 0, 3, 2)])
def test_tp2vec_equals_the_published_curl_formulas(kit3d, m, k, jpsi, jchi):
    mpmath.mp.dps = 40
    kit = kit3d
    nr, npts, nz = kit.p.nr, kit.p.np, kit.p.nz
    nph = npts // 2
    cpsi, cchi = 0.8 - 0.3j, -0.4 + 0.9j
    psi, chi = mo.scalar_init(kit, "FFF"), mo.scalar_init(kit, "FFF")
    psi.e[jpsi, m, k] = cpsi
    chi.e[jchi, m, k] = cchi
    vr, vp, vz = mo.tp2vec(psi, chi, kit)

    nodes = list(range(0, nr, 3))
    kap = float(kit.ak[k])                        # axial wavenumber 2 pi k' / zlen (k' signed)
    P = _mode_and_derivatives(kit, m + jpsi, m, nodes)
    C = _mode_and_derivatives(kit, m + jchi, m, nodes)
    phi = 2 * np.pi * np.arange(npts) / npts
    z = kit.p.zlen * np.arange(nz) / nz
    E = np.exp(1j * (m * phi[:, None] + kap * z[None, :]))            # [p, l]
    amp_r = np.array([complex(1j * m * cpsi * p0 / r + 1j * kap * cchi * c1) for (r, p0, _, _), (_, _, c1, _) in zip(P, C)])
    amp_p = np.array([complex(-cpsi * p1 - m * kap * cchi * c0 / r) for (r, _, p1, _), (_, c0, _, _) in zip(P, C)])
    amp_z = np.array([complex(-cchi * (c2 + c1 / r - mpmath.mpf(m * m) * c0 / (r * r))) for (r, c0, c1, c2) in C])
    for got_s, amp, name in ((vr, amp_r, "vr"), (vp, amp_p, "vp"), (vz, amp_z, "vz")):
        assert got_s.space == "PPP"
        got = np.empty((nr, npts, nz))
        got[:, 0::2, :] = got_s.e[:nr, :nph, :nz].real
        got[:, 1::2, :] = got_s.e[:nr, :nph, :nz].imag
        want = 2.0 * (amp[:, None, None] * E[None, :, :]).real          # the m < 0 partner is the conjugate
        scale = np.max(np.abs(want))
        assert scale > 0
        assert np.max(np.abs(got[nodes] - want)) / scale < 5e-11, name


def test_vec2tp_inverts_the_pinned_tp2vec_on_a_3d_field(kit3d):
    """vec2tp (ops:1308-1453: psi = -del_perp^-2 (curl V)_z, chi = -del_perp^-2 V_z, docs/tutorial/vector_field.md:50-52)
    must return the (psi, chi) that tp2vec -- pinned term by term above -- was given, for a field with azimuthal AND axial
    structure (the tutorial's q-vortex only exercises m = 0, k = 0).  The gauge is the reference's: both scalars vanish
    at infinity (zeroat1), which fixes the otherwise free P_{L_0}^0 component."""
    kit = kit3d
    rng = np.random.default_rng(5)

    def smooth_random():
        s = mo.scalar_init(kit, "FFF")
        for m in range(9):
            for k in (0, 1, 3):
                if m == 0 and k == 0:
                    continue
                s.e[:10, m, k] = (rng.standard_normal(10) + 1j * rng.standard_normal(10)) * np.exp(-np.arange(10) / 3)
        s.e[:, 0, 3] = np.conj(s.e[:, 0, 1])      # a real field: the m = 0 column is Hermitian in k
        mo.chop(s, kit)
        mo.zeroat1(s, kit)
        return s

    psi, chi = smooth_random(), smooth_random()
    vr, vp, vz = mo.tp2vec(psi, chi, kit)
    psi2, chi2 = mo.scalar_init(kit, "FFF"), mo.scalar_init(kit, "FFF")
    mo.vec2tp(vr, vp, vz, psi2, chi2, kit)
    assert np.linalg.norm(psi2.e - psi.e) / np.linalg.norm(psi.e) < 1e-12
    assert np.linalg.norm(chi2.e - chi.e) / np.linalg.norm(chi.e) < 5e-12
    assert abs(psi2.ln) < 1e-13 and abs(chi2.ln) < 1e-13


@pytest.mark.parametrize("power,alpha,beta", [(4, 50.0, -3.0), (6, -200.0, 2.0), (8, 1.0e4, 0.5)])
def test_ihelmp_inverts_helmp_built_from_the_pinned_del2(kit3d, power, alpha, beta):
    """helmp = del^p + beta del^2 + alpha applied as repeated del2 (pinned against Legendre's equation above); ihelmp must
    undo it: LAPACK's zgbtrf/zgbtrs on the band product the oracle assembles (ops:958-965) against the operator applied
    term by term.  Signs as the integrators use them (definite operators).  The axial Nyquist plane is left empty: with
    nzchop = nz/2 + 1 the reference's two k loops both visit it (kept, k_ranges), so there the two are not inverses."""
    kit = kit3d
    rng = np.random.default_rng(7)
    s = mo.scalar_init(kit, "FFF")
    s.e[:] = (rng.standard_normal(s.e.shape) + 1j * rng.standard_normal(s.e.shape)) \
        * np.exp(-np.arange(s.e.shape[0]) / 4.0)[:, None, None]
    s.e[:, :, kit.p.nz // 2] = 0.0
    mo.chop(s, kit)
    x = s.copy()
    mo.ihelmp(x, power, alpha, beta, kit)
    y = x.copy()
    mo.helmp(y, power, alpha, beta, kit)
    mo.chop(y, kit)
    assert np.linalg.norm(y.e - s.e) / np.linalg.norm(s.e) < 1e-11


def test_ihelm_inverts_del2_plus_alpha(kit3d):
    kit = kit3d
    rng = np.random.default_rng(8)
    s = mo.scalar_init(kit, "FFF")
    s.e[:] = (rng.standard_normal(s.e.shape) + 1j * rng.standard_normal(s.e.shape)) \
        * np.exp(-np.arange(s.e.shape[0]) / 4.0)[:, None, None]
    s.e[:, :, kit.p.nz // 2] = 0.0
    mo.chop(s, kit)
    alpha = -40.0                                  # -2 / (dt visc) of abcn is negative: del^2 + alpha is definite
    x = s.copy()
    mo.ihelm(x, alpha, kit)
    y = x.copy()
    mo.del2(y, kit)
    y.e = y.e + alpha * x.e
    mo.chop(y, kit)
    assert np.linalg.norm(y.e - s.e) / np.linalg.norm(s.e) < 1e-12


# ---- the semi-implicit integrators from their published discretisation ----------------------------------------------
# docs/tutorial/time_integration.md:160-169, 381-390:  L = VISC del^2 - HYPERVISC (-del^2)^(HYPERPOW/2),
#   febe:  w_{k+1} = w_k + dt A(w_k) + dt L(w_{k+1})
#   abcn:  w_{k+1} = w_k + dt [1.5 A(w_k) - 0.5 A(w_{k-1})] + dt [0.5 L(w_{k+1}) + 0.5 L(w_k)]
# The residual of those two equations is evaluated with L applied as repeated del2 (pinned above), not with the band
# products and LAPACK factorisations the integrators themselves run.

def _kit_with_viscosity(kit, visc, hyperpow, hypervisc):
    from dataclasses import replace
    p = replace(kit.p, visc=visc, hyperpow=hyperpow, hypervisc=hypervisc)
    tables = dict(x=kit.x, w=kit.w, lognorm=kit.lognorm, pf=kit.pf, at0=kit.at0, at1=kit.at1)
    return mo.kit_init(p, tables=tables)


def _apply_L(s, kit):
    """VISC del^2 s - HYPERVISC (-del^2)^(HYPERPOW/2) s, term by term."""
    p = kit.p
    a = s.copy()
    mo.del2(a, kit)
    out = p.visc * a.e
    if p.hyperpow:
        b = s.copy()
        for _ in range(p.hyperpow // 2):
            mo.del2(b, kit)
            b.e = -b.e
        out = out - p.hypervisc * b.e
    return out


def _smooth_field(kit, seed):
    rng = np.random.default_rng(seed)
    s = mo.scalar_init(kit, "FFF")
    s.e[:] = (rng.standard_normal(s.e.shape) + 1j * rng.standard_normal(s.e.shape)) \
        * np.exp(-np.arange(s.e.shape[0]) / 3.0)[:, None, None]
    s.e[:, :, kit.p.nz // 2] = 0.0                 # the doubly visited Nyquist plane, see above
    mo.chop(s, kit)
    return s


@pytest.mark.parametrize("visc,hyperpow,hypervisc", [(1.0e-2, 0, 0.0), (1.0e-3, 4, 1.0e-5), (1.0e-3, 8, 1.0e-9)])
def test_febe_and_abcn_satisfy_the_published_discretisation(kit3d, visc, hyperpow, hypervisc):
    kit = _kit_with_viscosity(kit3d, visc, hyperpow, hypervisc)
    dt = 0.05
    w0, nl, nl_p = _smooth_field(kit, 21), _smooth_field(kit, 22), _smooth_field(kit, 23)

    w1 = w0.copy()
    mo.febe(w1, nl, dt, kit)
    res = w1.e - w0.e - dt * nl.e - dt * _apply_L(w1, kit)
    r = mo.Scalar(e=np.asfortranarray(res), space="FFF")
    mo.chop(r, kit)
    assert np.linalg.norm(r.e) / np.linalg.norm(w1.e) < 1e-11

    w1 = w0.copy()
    w_prev, nl_prev = w0.copy(), nl_p.copy()
    Lw0 = _apply_L(w0, kit)
    mo.abcn(w1, w_prev, nl, nl_prev, dt, kit)
    res = w1.e - w0.e - dt * (1.5 * nl.e - 0.5 * nl_p.e) - dt * (0.5 * _apply_L(w1, kit) + 0.5 * Lw0)
    r = mo.Scalar(e=np.asfortranarray(res), space="FFF")
    mo.chop(r, kit)
    assert np.linalg.norm(r.e) / np.linalg.norm(w1.e) < 1e-11
    # ops:1256: the previous-step slots receive the NEW field and the current nonlinear term
    assert np.array_equal(w_prev.e, w1.e) and np.array_equal(nl_prev.e, nl.e)


def test_end_point_tables_are_the_closed_forms(kit3d):
    """at0 / at1 (sinit:136-150: the m = 0 functions at x = -1 + 1e-15 and 1 - 1e-15) against P_n(-1) = (-1)^n,
    P_n(1) = 1 with the orthonormal weight sqrt((2n+1)/2); the 1e-15 offset moves them by n(n+1)/2 * 1e-15."""
    n = np.arange(kit3d.p.nrchop)
    norm = np.sqrt((2.0 * n + 1.0) / 2.0)
    assert np.max(np.abs(kit3d.at1 / norm - 1.0)) < 2e-12
    assert np.max(np.abs(kit3d.at0 / norm - (-1.0) ** n)) < 2e-12
    # calcat0 / calcat1 of a field = the synthesis of its m = 0 column at those points
    s = mo.scalar_init(kit3d, "FFF")
    s.e[:6, 0, 0] = [1.0, -0.5, 0.25, 2.0, 0.0, -1.0]
    s.e[:6, 0, 1] = [0.5j, 1.0, 0.0, -1.0j, 0.25, 0.0]
    c0, c1 = mo.calcat0(s, kit3d), mo.calcat1(s, kit3d)
    for k in (0, 1):
        col = s.e[:6, 0, k]
        assert abs(c0[k] - np.sum(col * norm[:6] * (-1.0) ** n[:6])) < 1e-12
        assert abs(c1[k] - np.sum(col * norm[:6])) < 1e-12
    mo.zeroat1(s, kit3d)
    assert np.max(np.abs(mo.calcat1(s, kit3d))) < 1e-14


def test_delsqp_is_the_weighted_horizontal_laplacian(kit3d):
    """docs/tutorial/operation.md:300-302: ((L^2 + r^2) / (2 L^2))^2 del_perp^2 P_{L_n}^m = -n(n+1)/L^2 P_{L_n}^m.  delsqp's
    diagonal factors against the pinned del2h evaluated on the grid and weighted there, for low-degree content."""
    kit = kit3d
    rng = np.random.default_rng(31)
    s = mo.scalar_init(kit, "FFF")
    for m in range(6):
        s.e[:8, m, :2] = rng.standard_normal((8, 2)) + 1j * rng.standard_normal((8, 2))
    s.e[0, 0, :] = 0.0                              # n = 0: the logarithmic term's slot (ops:360)
    a, b = s.copy(), s.copy()
    mo.delsqp(a, kit)
    mo.del2h(b, kit)
    a.space = "FFP"
    b.space = "FFP"
    mo.rtrans_backward(a, kit)
    mo.rtrans_backward(b, kit)
    nr = kit.p.nr
    wgt = ((kit.p.ell ** 2 + kit.r ** 2) / (2.0 * kit.p.ell ** 2)) ** 2
    # compared unweighted: the weight reaches 1e6 at the outermost node and would only amplify del2h's rounding
    lhs = a.e[:nr, :6, :2] / wgt[:, None, None]
    rhs = b.e[:nr, :6, :2]
    assert np.max(np.abs(lhs - rhs)) / np.max(np.abs(rhs)) < 1e-12
    inv = s.copy()
    mo.delsqp(inv, kit)
    mo.idelsqp(inv, kit)
    assert np.linalg.norm(inv.e - s.e) / np.linalg.norm(s.e) < 1e-14
