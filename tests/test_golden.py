"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle with fixed seeds).

CPU: the oracle and the C++ binary128 table builder reproduce them (so an edit of the oracle that changes numbers
is caught).  GPU: the CUDA path, fed the golden tables, reproduces them through the C ABI."""
import os

import numpy as np
import pytest

import mlegs_b200 as mb
from mlegs_b200 import vortex
from oracle import mlegs_oracle as mo
from helpers import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1.0e-12
CFG = {
    "gate2d": dict(nr=32, np=48, nz=1, nrchop=32, npchop=25, nzchop=1, ell=1.0, visc=5e-3, hyperpow=0, hypervisc=0.0),
    "gate3d": dict(nr=32, np=16, nz=8, nrchop=32, npchop=9, nzchop=5, ell=4.0, visc=1e-4, hyperpow=8, hypervisc=5e-7),
}


def _load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


def _params(name):
    c = CFG[name]
    return mb.make_params(c["nr"], c["np"], c["nz"], c["nrchop"], c["npchop"], c["nzchop"], ell=c["ell"],
                          zlen=2 * np.pi, visc=c["visc"], hyperpow=c["hyperpow"], hypervisc=c["hypervisc"])


@pytest.mark.parametrize("name", list(CFG))
def test_host_tables_match_golden(name):
    g = _load(name)
    kit = mb.TfmKit.build_tables(_params(name))     # C++ binary128 recurrence, no GPU
    assert np.array_equal(kit.x, g["x"]) and np.array_equal(kit.w, g["w"])
    assert np.array_equal(kit.lognorm, g["lognorm"])
    assert np.max(np.abs(kit.pf - g["pf"])) <= 4 * np.finfo(float).eps * np.max(np.abs(g["pf"]))
    assert np.allclose(kit.at0, g["at0"], rtol=1e-13, atol=0) and np.allclose(kit.at1, g["at1"], rtol=1e-13, atol=0)


@pytest.mark.parametrize("name", list(CFG))
def test_oracle_reproduces_golden(name):
    g = _load(name)
    c = CFG[name]
    kit = mo.kit_init(mo.Params(zlen=2.0 * np.pi, **c),
                      tables={k: g[k] for k in ("x", "w", "lognorm", "pf", "at0", "at1")})
    s = mo.Scalar(e=g["fff0"].copy(order="F"), space="FFF", ln=0.25)
    mo.trans(s, "PPP", kit)
    assert rel_l2(s.e, g["ppp_ln025"]) < 1e-14
    mo.trans(s, "FFF", kit)
    assert rel_l2(s.e, g["fff_roundtrip"]) < 1e-14
    for op in ("del2", "del2h", "xxdx", "delsqp"):
        t = mo.Scalar(e=g["fff0"].copy(order="F"), space="FFF", ln=0.25)
        getattr(mo, op)(t, kit)
        assert rel_l2(t.e, g[op]) < 1e-14 and t.ln == float(g[op + "_ln"])
    t = mo.Scalar(e=g["fff0"].copy(order="F"), space="FFF")
    mo.ihelm(t, -7.5, kit)
    assert rel_l2(t.e, g["ihelm_m7p5"]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CFG))
def test_cuda_reproduces_golden(name):
    g = _load(name)
    p = _params(name)
    kit = mb.TfmKit.build_tables(p)
    for k in ("x", "w", "lognorm", "pf", "at0", "at1"):     # the golden tables, bit for bit
        getattr(kit, k)[...] = g[k]
    kit.upload()

    def fresh(ln=0.0):
        s = mb.Scalar("FFF").upload(g["fff0"])
        s.ln = ln
        return s

    s = fresh(0.25)
    mb.trans(s, "PPP")
    assert rel_l2(s.download(), g["ppp_ln025"]) < TOL
    mb.trans(s, "FFF")
    assert rel_l2(s.download(), g["fff_roundtrip"]) < TOL
    for op in ("del2", "del2h", "xxdx", "delsqp"):
        t = fresh(0.25)
        getattr(mb, op)(t)
        assert rel_l2(t.download(), g[op]) < TOL, op
        assert abs(t.ln - float(g[op + "_ln"])) <= 1e-15 * max(1.0, abs(float(g[op + "_ln"])))
    t = fresh()
    mb.ihelm(t, -7.5)
    assert rel_l2(t.download(), g["ihelm_m7p5"]) < TOL
    t = fresh()
    mb.idel2(t)
    assert rel_l2(t.download(), g["idel2"]) < TOL
    assert abs(t.ln - float(g["idel2_ln"])) <= 1e-12 * max(1.0, abs(float(g["idel2_ln"])))
    t = fresh()
    gain = mb.svv_filter(t, 0.3)
    assert abs(gain - float(g["svv_gain"])) <= 1e-13 and rel_l2(t.download(), g["svv"]) < TOL
    if "ihelmp" in g:
        c = CFG[name]
        t = fresh()
        mb.ihelmp(t, c["hyperpow"], -2.0 / (1e-2 * c["hypervisc"] * -1.0), c["visc"] / (c["hypervisc"] * -1.0))
        assert rel_l2(t.download(), g["ihelmp"]) < TOL
    if "step2_psi" in g:
        dt = 1e-2
        psi, chi = vortex.qvort_dist_tp(kit)
        uz = vortex.uniform_z_fld(kit)
        assert rel_l2(psi.download(), g["qvort_psi0"]) < TOL and rel_l2(chi.download(), g["qvort_chi0"]) < TOL
        st = vortex.bootstrap(kit, dt, psi, chi, uz)
        assert rel_l2(st.psi.download(), g["boot_psi"]) < TOL and rel_l2(st.chi.download(), g["boot_chi"]) < TOL
        for _ in range(2):
            vortex.step(st, dt)
        assert rel_l2(st.psi.download(), g["step2_psi"]) < TOL and rel_l2(st.chi.download(), g["step2_chi"]) < TOL
        assert abs(st.psi.ln - float(g["step2_psi_ln"])) <= 1e-10 * max(1.0, abs(float(g["step2_psi_ln"])))
