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
