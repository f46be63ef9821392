"""Time-step parity at a BASELINE size on one GPU: the q-vortex run of BASELINE.json configs[2] (input.params physics:
hyperpow 8, SVV on, de-aliasing, ran_noise = 0) at NR = NP = NZ = 128 -- Richardson bootstrap + 3 ABCN steps
(apps/vortical_flow_3d.f90:116-180) -- against the oracle, 1e-12 relative L2 on psi and chi after the bootstrap and
after every step.  The oracle run (about a minute on 16 host cores) happens in a process of its own
(tests/oracle_vortex.py) while the device run is prepared."""
import os
import subprocess
import sys

import numpy as np
import pytest

import mlegs_b200 as mb
from mlegs_b200 import vortex
from helpers import rel_l2
from oracle_vortex import DT, qvortex_params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1.0e-12
N, STEPS = 128, 3


def test_qvortex_time_steps_128(tmp_path):
    out = str(tmp_path / "qv128.npz")
    proc = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "oracle_vortex.py"), "--n", str(N), "--steps",
                             str(STEPS), "--out", out], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                            text=True)
    kit = mb.TfmKit.init(qvortex_params(N))
    psi, chi = vortex.qvort_dist_tp(kit, q=1.0)
    uz = vortex.uniform_z_fld(kit, b=-0.5)
    got = {"psi_ic": psi.download(), "chi_ic": chi.download()}
    st = vortex.bootstrap(kit, DT, psi, chi, uz)
    meta = []
    for it in range(STEPS + 1):
        got[f"psi_{it}"], got[f"chi_{it}"] = st.psi.download(), st.chi.download()
        meta.append([st.psi.ln, st.chi.ln, st.gain_psi, st.gain_chi])
        if it < STEPS:
            vortex.step(st, DT)
    log, _ = proc.communicate(timeout=1500)
    assert proc.returncode == 0, log[-3000:]
    want = np.load(out)
    errs = {}
    for key in got:
        errs[key] = rel_l2(got[key], want[key])
    print("q-vortex 128^3 rel-L2 vs oracle:", {k: f"{v:.2e}" for k, v in errs.items()})
    for key, err in errs.items():
        assert err < TOL, (key, err, errs)
    wm = want["meta"]
    for it in range(STEPS + 1):
        for q in range(2):
            assert abs(meta[it][q] - wm[it][q]) <= TOL * max(1.0, abs(wm[it][q])), ("ln", it, q)
            assert abs(meta[it][2 + q] - wm[it][2 + q]) <= TOL, ("svv gain", it, q, meta[it], wm[it])
    # the state is not trivially small: the SVV controller and the nonlinear term are active
    assert np.linalg.norm(want["psi_3"] - want["psi_0"]) > 1e-6 * np.linalg.norm(want["psi_0"])
