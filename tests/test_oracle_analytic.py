"""Pins the oracle against the analytic known-answers of the reference's tutorials
(SURVEY.md section 4.3).  The reference asserts no numbers itself ("parity unpinned")."""
import math

import numpy as np
import pytest

from oracle import mlegs_oracle as mo


@pytest.fixture(scope="module")
def kit2d():
    # tools/validate_tutorials.py:255-269 / input_2d.params: NR=32, NP=48, NZ=1, L=1
    p = mo.Params(nr=32, np=48, nz=1, nrchop=32, npchop=25, nzchop=1, ell=1.0, zlen=1.0,
                  visc=5.0e-3, hyperpow=0, hypervisc=0.0)
    return mo.kit_init(p)


def _phys_grid(kit):
    r = kit.r[:, None]
    nph = kit.p.np // 2
    phi = 2.0 * mo.PI / kit.p.np * np.arange(kit.p.np)
    return r, phi[0::2][None, :nph], phi[1::2][None, :nph]


def _two_mode_field(kit):
    s = mo.scalar_init(kit, "FFF")
    s.e[1, 1, 0] = 1.0   # e(2,2,1)
    s.e[1, 2, 0] = 1.0   # e(2,3,1)
    return s


def _analytic_s(r, phi):
    # docs/tutorial/transformation.md:215-221
    return 2.0 * (math.sqrt(5.0 / 12.0) * (-6.0 * r * (r ** 2 - 1.0) / (r ** 2 + 1.0) ** 2) * np.cos(phi)
                  + math.sqrt(7.0 / 240.0) * (60.0 * r ** 2 * (r ** 2 - 1.0) / (r ** 2 + 1.0) ** 3) * np.cos(2 * phi))


def test_tables_match_analytic_functions(kit2d):
    # P_{L_2}^1 and P_{L_3}^2 in closed form (same source), on the first nr/2 nodes
    r = kit2d.r[: kit2d.p.nr // 2]
    f1 = math.sqrt(5.0 / 12.0) * (-6.0 * r * (r ** 2 - 1.0) / (r ** 2 + 1.0) ** 2)
    f2 = math.sqrt(7.0 / 240.0) * (60.0 * r ** 2 * (r ** 2 - 1.0) / (r ** 2 + 1.0) ** 3)
    assert np.max(np.abs(kit2d.pf[:, 1, 1] - f1)) < 5e-15
    assert np.max(np.abs(kit2d.pf[:, 1, 2] - f2)) < 5e-15


def test_table_orthonormal(kit2d):
    nrh = kit2d.p.nr // 2
    for m in (0, 3, 24):
        nn = int(kit2d.chops[m])
        full = np.zeros((kit2d.p.nr, nn))
        par = (-1.0) ** np.arange(nn)
        full[:nrh] = kit2d.pf[:, :nn, m]
        full[::-1][:nrh] = kit2d.pf[:, :nn, m] * par[None, :]
        gram = full.T @ (kit2d.w[:, None] * full)
        # the quadrature is exact only while the integrand degree stays < 2 nr
        ok = [(i, j) for i in range(nn) for j in range(nn) if (2 * m + i + j) < 2 * kit2d.p.nr - 1]
        err = max(abs(gram[i, j] - (1.0 if i == j else 0.0)) for i, j in ok)
        assert err < 2e-13


def test_backward_transform_known_answer(kit2d):
    s = _two_mode_field(kit2d)
    mo.trans(s, "PPP", kit2d)
    r, phi_re, phi_im = _phys_grid(kit2d)
    nr, nph = kit2d.p.nr, kit2d.p.np // 2
    assert np.max(np.abs(s.e[:nr, :nph, 0].real - _analytic_s(r, phi_re))) < 1e-13
    assert np.max(np.abs(s.e[:nr, :nph, 0].imag - _analytic_s(r, phi_im))) < 1e-13


def test_forward_of_backward_is_identity(kit2d):
    s = _two_mode_field(kit2d)
    ref = s.e.copy()
    mo.trans(s, "PPP", kit2d)
    mo.trans(s, "FFF", kit2d)
    assert np.linalg.norm(s.e - ref) / np.linalg.norm(ref) < 1e-13


def test_laplacian_known_answer(kit2d):
    # docs/tutorial/operation.md:150
    s = _two_mode_field(kit2d)
    mo.del2(s, kit2d)
    mo.trans(s, "PPP", kit2d)
    r, phi_re, _ = _phys_grid(kit2d)
    exact = 48.0 * math.sqrt(15.0) * r * (r ** 2 - 1.0) * ((r ** 2 + 1.0) * np.cos(phi_re)
                                                           - 2.0 * math.sqrt(7.0) * r * np.cos(2 * phi_re)) / (r ** 2 + 1.0) ** 5
    nr, nph = kit2d.p.nr, kit2d.p.np // 2
    assert np.max(np.abs(s.e[:nr, :nph, 0].real - exact)) < 1e-11


def test_inverse_laplacian_roundtrip(kit2d):
    # apps/inverse_laplacian.f90: idel2(del2(s)) == s
    s = _two_mode_field(kit2d)
    ref = s.e.copy()
    mo.del2(s, kit2d)
    mo.idel2_proln(s, kit2d)
    assert np.linalg.norm(s.e - ref) / np.linalg.norm(ref) < 1e-12


def test_gaussian_vortex_diffusion_orders():
    # apps/time_integration_second.f90:3-4,394-401: w(r,t) = exp(-r^2/(2(1+2 nu t)))/(2 pi (1+2 nu t));
    # FEBE is first order, ABCN second (docs/tutorial/time_integration.md:54)
    p = mo.Params(nr=32, np=16, nz=1, nrchop=32, npchop=9, nzchop=1, ell=2.0, zlen=1.0,
                  visc=1.0e-1, hyperpow=0, hypervisc=0.0, is_svv=False)
    kit = mo.kit_init(p)

    def initial():
        s = mo.scalar_init(kit, "PPP")
        s.e[: p.nr, : p.np // 2, 0] = (np.exp(-kit.r ** 2 / 2.0) / (2 * mo.PI))[:, None] * (1 + 1j)
        mo.trans(s, "FFF", kit)
        return s

    def exact0(t):
        return 1.0 / (2 * mo.PI * (1 + 2 * p.visc * t))

    def run_febe(dt, nsteps):
        s = initial()
        zero = mo.scalar_init(kit, "FFF")
        for _ in range(nsteps):
            mo.febe(s, zero, dt, kit)
        return abs(mo.calcat0(s, kit)[0].real - exact0(dt * nsteps))

    def run_abcn(dt, nsteps):
        s = initial()
        zero = mo.scalar_init(kit, "FFF")
        sp, nlp = s.copy(), zero.copy()
        for _ in range(nsteps):
            mo.abcn(s, sp, zero, nlp, dt, kit)
        return abs(mo.calcat0(s, kit)[0].real - exact0(dt * nsteps))

    e1, e2 = run_febe(0.1, 10), run_febe(0.05, 20)
    assert 1.7 < e1 / e2 < 2.3
    c1, c2 = run_abcn(0.1, 10), run_abcn(0.05, 20)
    assert 3.5 < c1 / c2 < 4.5
    assert c1 < e1


def test_qvortex_tp_roundtrip():
    # docs/tutorial/vector_field.md:88; apps/vecfld_reconstruction.f90, tp_project.f90:
    # psi, chi of a q-vortex -> V = (0, (1-exp(-r^2))/r, exp(-r^2)/q); vec2tp(tp2vec(psi,chi)) == (psi,chi)
    p = mo.Params(nr=48, np=8, nz=4, nrchop=48, npchop=5, nzchop=3, ell=3.0, zlen=2 * mo.PI,
                  visc=1e-3, hyperpow=0, hypervisc=0.0, is_svv=False)
    kit = mo.kit_init(p)
    q = 1.0
    fields = []
    for amp in (2.0, 1.0 / q):
        s = mo.scalar_init(kit, "PPP")
        s.e[: p.nr, : p.np // 2, : p.nz] = ((-np.exp(-kit.r ** 2) * amp / (1.0 - kit.x) ** 2) * (1 + 1j))[:, None, None]
        mo.trans(s, "FFF", kit)
        mo.idelsqp(s, kit)
        mo.zeroat1(s, kit)
        fields.append(s)
    psi, chi = fields
    vr, vp, vz = mo.tp2vec(psi, chi, kit)
    r = kit.r
    nr = p.nr
    vphi = (1.0 - np.exp(-r ** 2)) / r
    vzz = np.exp(-r ** 2) / q
    assert np.max(np.abs(vp.e[:nr, 0, 0].real - vphi)) < 1e-12
    assert np.max(np.abs(vp.e[:nr, 0, 0].imag - vphi)) < 1e-12
    assert np.max(np.abs(vz.e[:nr, 0, 0].real - vzz)) < 1e-12
    assert np.max(np.abs(vr.e[:nr, : p.np // 2, : p.nz])) < 1e-12
    assert abs(psi.ln + 0.5) < 1e-13 and abs(chi.ln + 0.25 / q) < 1e-13
    psi2, chi2 = mo.scalar_init(kit, "FFF"), mo.scalar_init(kit, "FFF")
    mo.vec2tp(vr, vp, vz, psi2, chi2, kit)
    assert np.linalg.norm(psi2.e - psi.e) / np.linalg.norm(psi.e) < 1e-9
    assert np.linalg.norm(chi2.e - chi.e) / np.linalg.norm(chi.e) < 1e-9


def test_forked_column_sweeps_are_bit_identical():
    """mo.set_workers(n) runs the per-(m,k) sweeps of the band operators and solves on forked processes, one azimuthal
    column per task; the arithmetic per column is the same code, so results must equal the serial sweep bit for bit."""
    import mlegs_b200 as mb
    from helpers import oracle_params, random_fff
    p = mb.make_params(24, 12, 8, 24, 7, 5, ell=3.0, zlen=2 * np.pi, visc=1e-3, hyperpow=4, hypervisc=1e-6)
    kit = mb.TfmKit.build_tables(p)
    ok = mo.kit_init(oracle_params(p), tables=kit.tables())
    e = random_fff(ok, seed=3)

    def run_all():
        out = []
        for fn in (lambda s: mo.del2(s, ok), lambda s: mo.xxdx(s, ok), lambda s: mo.ihelm(s, -50.0, ok),
                   lambda s: mo.ihelmp(s, 4, -1e4, 3.0, ok), lambda s: mo.idel2_proln(s, ok),
                   lambda s: mo.idel2_preln(s, ok, 0.3)):
            s = mo.Scalar(e=e.copy(order="F"), space="FFF", ln=0.2)
            fn(s)
            out.append((s.e.copy(), s.ln))
        return out

    mo.set_workers(1)
    serial = run_all()
    mo.set_workers(3)
    try:
        forked = run_all()
    finally:
        mo.set_workers(1)
    for (a, la), (b, lb) in zip(serial, forked):
        assert np.array_equal(a, b) and la == lb


def test_smooth_and_fftreat_properties():
    """smooth (ops:2149-2175) keeps the sum of a tail whose ends it does not touch, and reproduces the reference's
    hand-computable 4-point case; fftreat (ops:1002-1063) leaves the far-field value at r -> infinity at zero."""
    a = np.array([1.0, 2.0, 4.0, 8.0], dtype=np.complex128)
    # i = 2: f = (1 + 2/3)/2 = 5/6 -> (1/12, 5/6, 1/12) * 2;  i = 3: f = (1 + 1/3)/2 = 2/3 -> (1/6, 2/3, 1/6) * 4;
    # first point kept, last point spread (.2, .3, .5) over the last three
    want = np.array([1.0 + 2.0 / 12.0, 8.0 * 0.2 + 2.0 * 5.0 / 6.0 + 4.0 / 6.0, 8.0 * 0.3 + 2.0 / 12.0 + 4.0 * 2.0 / 3.0,
                     8.0 * 0.5 + 4.0 / 6.0])
    assert np.allclose(mo.smooth(a), want, rtol=0, atol=1e-15)
    assert abs(mo.smooth(a).sum() - a.sum()) < 1e-14          # every weight triple sums to one
    import mlegs_b200 as mb
    from helpers import oracle_params, random_fff
    p = mb.make_params(32, 16, 8, 32, 9, 5, ell=4.0, zlen=2 * np.pi)
    kit = mb.TfmKit.build_tables(p)
    ok = mo.kit_init(oracle_params(p), tables=kit.tables())
    s = mo.Scalar(e=random_fff(ok, seed=2, decay=1.0), space="FFF", ln=0.2)
    mo.fftreat(s, ok)
    assert s.space == "FFF" and s.ln == 0.2
    assert np.max(np.abs(mo.calcat1(s, ok))) < 1e-12 * np.max(np.abs(s.e))
