#!/usr/bin/env python
"""Generates tests/golden/*.npz from the oracle (oracle/mlegs_oracle.py) with fixed seeds.

The reference itself cannot be run (Fortran + MPI, no compiler in the image or on the GPU box), so these
vectors pin the ORACLE -- which is in turn pinned by the tutorials' analytic answers
(tests/test_oracle_analytic.py) -- and give the CUDA path a fixed target that does not move when the oracle
is edited.  Regenerate with `python tests/golden/make_golden.py` (needs mpmath for the 50-digit tables).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import mlegs_oracle as mo  # noqa: E402
from helpers import random_fff  # noqa: E402

CASES = {
    # tools/validate_tutorials.py:255-269 (2-D gate) and :222-238 (3-D gate, input.params physics)
    "gate2d": dict(nr=32, np=48, nz=1, nrchop=32, npchop=25, nzchop=1, ell=1.0, visc=5e-3, hyperpow=0, hypervisc=0.0),
    "gate3d": dict(nr=32, np=16, nz=8, nrchop=32, npchop=9, nzchop=5, ell=4.0, visc=1e-4, hyperpow=8, hypervisc=5e-7),
}


def make(name, cfg):
    p = mo.Params(zlen=2.0 * np.pi, **cfg)
    kit = mo.kit_init(p)
    out = dict(x=kit.x, w=kit.w, lognorm=kit.lognorm, pf=kit.pf, at0=kit.at0, at1=kit.at1)
    e0 = random_fff(kit, seed=11)
    out["fff0"] = e0
    s = mo.Scalar(e=e0.copy(order="F"), space="FFF", ln=0.25)
    mo.trans(s, "PPP", kit)
    out["ppp_ln025"] = s.e.copy(order="F")
    mo.trans(s, "FFF", kit)
    out["fff_roundtrip"] = s.e.copy(order="F")
    for op in ("del2", "del2h", "xxdx", "delsqp"):
        t = mo.Scalar(e=e0.copy(order="F"), space="FFF", ln=0.25)
        getattr(mo, op)(t, kit)
        out[op] = t.e.copy(order="F")
        out[op + "_ln"] = np.float64(t.ln)
    t = mo.Scalar(e=e0.copy(order="F"), space="FFF")
    mo.ihelm(t, -7.5, kit)
    out["ihelm_m7p5"] = t.e.copy(order="F")
    t = mo.Scalar(e=e0.copy(order="F"), space="FFF")
    mo.idel2_proln(t, kit)
    out["idel2"] = t.e.copy(order="F")
    out["idel2_ln"] = np.float64(t.ln)
    if cfg["hyperpow"]:
        t = mo.Scalar(e=e0.copy(order="F"), space="FFF")
        mo.ihelmp(t, cfg["hyperpow"], -2.0 / (1e-2 * cfg["hypervisc"] * -1.0), cfg["visc"] / (cfg["hypervisc"] * -1.0), kit)
        out["ihelmp"] = t.e.copy(order="F")
    t = mo.Scalar(e=e0.copy(order="F"), space="FFF")
    out["svv_gain"] = np.float64(mo.svv_filter(t, kit, 0.3))
    out["svv"] = t.e.copy(order="F")
    if cfg["nz"] > 1:
        dt = 1e-2
        psi, chi = mo.qvort_dist_tp(kit)
        uz = mo.uniform_z_fld(kit)
        out["qvort_psi0"], out["qvort_chi0"] = psi.e.copy(order="F"), chi.e.copy(order="F")
        st = mo.vortex_bootstrap(kit, dt, psi, chi, uz)
        out["boot_psi"], out["boot_chi"] = st.psi.e.copy(order="F"), st.chi.e.copy(order="F")
        for it in range(2):
            mo.vortex_step(st, kit, dt)
        out["step2_psi"], out["step2_chi"] = st.psi.e.copy(order="F"), st.chi.e.copy(order="F")
        out["step2_psi_ln"] = np.float64(st.psi.ln)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", ()) for k, v in out.items()})


if __name__ == "__main__":
    for name, cfg in CASES.items():
        make(name, cfg)
