"""Oracle snapshots of the q-vortex run (BASELINE.json configs[2] physics: input.params values, two q-vortices at
x = -2, +2, ran_noise = 0): psi, chi after the Richardson bootstrap and after each ABCN step, computed by the NumPy
oracle (oracle/mlegs_oracle.py: vortex_bootstrap, vortex_step = apps/vortical_flow_3d.f90:116-180) in a process of its
own -- no CUDA here, so the per-column sweeps may fork (mo.set_workers).  Test infrastructure: the GPU parity tests at
128^3 (one GPU) and 64^3 (2/4/8 GPUs) launch this next to the device run and compare snapshot by snapshot.

    python tests/oracle_vortex.py --n 128 --steps 3 --workers 16 --out /tmp/qv128.npz
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DT = 1.0e-2


def qvortex_params(n: int, nz: int | None = None):
    """input.params:7-29 with NR = NP = n, NZ = nz (SURVEY.md section 8d inputs 3/4)."""
    import mlegs_b200 as mb
    nz = n if nz is None else nz
    return mb.make_params(n, n, nz, n, n // 2 + 1, nz // 2 + 1, ell=4.0, zlen=2 * np.pi, visc=1.0e-4, hyperpow=8,
                          hypervisc=5.0e-7, is_svv=True, svv_cutoff=0.75, svv_target=2.0e-2, svv_strength=0.12,
                          svv_relax=0.25)


def run(n: int, steps: int, workers: int, out: str, nz: int | None = None):
    import mlegs_b200 as mb            # host-only use: the binary128 table builder behind the C ABI (no GPU needed)
    from oracle import mlegs_oracle as mo
    from helpers import oracle_params
    p = qvortex_params(n, nz)
    kit = mb.TfmKit.build_tables(p)
    ok = mo.kit_init(oracle_params(p), tables=kit.tables())
    mo.set_workers(workers)
    t0 = time.time()
    psi, chi = mo.qvort_dist_tp(ok, q=1.0)
    uz = mo.uniform_z_fld(ok, b=-0.5)
    snaps = {"psi_ic": psi.e.copy(order="F"), "chi_ic": chi.e.copy(order="F")}
    st = mo.vortex_bootstrap(ok, DT, psi, chi, uz)
    meta = []
    for it in range(steps + 1):
        snaps[f"psi_{it}"] = st.psi.e.copy(order="F")
        snaps[f"chi_{it}"] = st.chi.e.copy(order="F")
        meta.append([st.psi.ln, st.chi.ln, st.gain_psi, st.gain_chi])
        print(f"oracle_vortex: snapshot {it} at {time.time() - t0:.1f} s", flush=True)
        if it < steps:
            mo.vortex_step(st, ok, DT)
    snaps["meta"] = np.array(meta)
    tmp = out + ".tmp.npz"
    np.savez(tmp, **snaps)
    os.replace(tmp, out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--nz", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--workers", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    run(a.n, a.steps, a.workers, a.out, a.nz or None)
