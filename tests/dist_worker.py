"""Multi-GPU parity worker (launched by tests/test_gpu_dist.py under torchrun, one rank per GPU).

Every rank builds the same oracle state (single process, global arrays) and checks its own slab of the
device result against the matching slab of the oracle: exchange (bit-exact), transforms, operators with
cross-rank reductions (SVV, far-field values, ln broadcast), vector operations and q-vortex time steps.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mlegs_b200 as mb  # noqa: E402
from mlegs_b200 import vortex as mv  # noqa: E402
from oracle import mlegs_oracle as mo  # noqa: E402
from helpers import oracle_kit, random_fff, rel_l2  # noqa: E402

TOL = 1.0e-12


def slab(s: mb.Scalar, glb: np.ndarray) -> np.ndarray:
    return glb[s.global_slices()]


def check(name, s: mb.Scalar, glb: np.ndarray, tol=TOL, exact=False):
    got, want = s.download(), slab(s, glb)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    if exact:
        ok = np.array_equal(got, want)
        err = 0.0 if ok else 1.0
    else:
        # global relative L2: local squared error / global squared norm, summed over ranks
        num, den = float(np.sum(np.abs(got - want) ** 2)), float(np.sum(np.abs(want) ** 2))
        num, den = mb.dist.allreduce([num, den])
        err = float(np.sqrt(num / den)) if den > 0 else float(np.sqrt(num))
        ok = err < tol
    if int(os.environ["RANK"]) == 0:
        print(f"  {name}: {'exact' if exact else f'rel-L2 {err:.2e}'}", flush=True)
    assert ok, (name, err)


def run_case(nr, np_, nz, nrc, npc, nzc, ell, hp, rank, world, steps):
    p = mb.make_params(nr, np_, nz, nrc, npc, nzc, ell=ell, zlen=2 * np.pi, visc=1e-4, hyperpow=hp,
                       hypervisc=(5e-7 if hp else 0.0))
    kit = mb.TfmKit.init(p, rank, world)
    mb.dist.attach()
    ok = oracle_kit(kit)
    if rank == 0:
        print(f"case nr={nr} np={np_} nz={nz} hyperpow={hp} on {world} ranks", flush=True)

    # ---- exchange: index-encoded fill (src/apps/assemble.f90:49), bit-exact data movement ----
    i, j, k = np.meshgrid(*[np.arange(n) for n in kit.glb_sz], indexing="ij")
    full = np.asfortranarray((i * 1e4 + j * 1e2 + k) + 1j * (k * 1e4 + i * 1e2 + j))
    s = mb.Scalar("PPP").upload_global(full)
    assert s.f.axis_comm[:] == [1, 0, 2]
    s.exchange(2, 1)                       # the chain of ops:185-208 / src/apps/assemble.f90
    assert s.f.axis_comm[:] == [0, 1, 2], s.f.axis_comm[:]
    check("exchange(2,1)", s, full, exact=True)
    s.exchange(1, 3)                       # comm_grps(2) is a single rank in the slab layout: re-labelling
    assert s.f.axis_comm[:] == [2, 1, 0], s.f.axis_comm[:]
    check("exchange(1,3)", s, full, exact=True)
    try:
        s.exchange(1, 2)                   # r is labelled distributed now: the reference aborts (dist:16-20)
        raise AssertionError("exchange(1,2) on (2,1,0) must fail")
    except mb.MlegsError as e:
        assert "non-distributed along the old dimension" in str(e)
    s.exchange(3, 1)
    s.exchange(1, 2)
    assert s.f.axis_comm[:] == [1, 0, 2], s.f.axis_comm[:]
    check("exchange(1,2)", s, full, exact=True)

    # ---- transforms ----
    e0 = random_fff(ok, seed=5)
    for ln in (0.0, 0.37):
        s = mb.Scalar("FFF").upload_global(e0)
        s.ln = ln
        so = mo.Scalar(e=e0.copy(order="F"), space="FFF", ln=ln)
        for sp in ("FFP", "PFP", "PPP", "PFP", "FFP", "FFF", "PPP", "FFF"):
            mb.trans(s, sp)
            mo.trans(so, sp, ok)
            check(f"trans -> {sp} (ln={ln})", s, so.e)

    # ---- batched transforms: one fused exchange for the whole group, bit-identical to one scalar at a time ----
    seeds = (11, 12, 13)
    group = [mb.Scalar("FFF").upload_global(random_fff(ok, seed=sd)) for sd in seeds]
    single = [mb.Scalar("FFF").upload_global(random_fff(ok, seed=sd)) for sd in seeds]
    for q, g1 in enumerate(group):
        g1.ln = single[q].ln = 0.1 * q
    for sp in ("PPP", "FFF", "PFP", "PPP", "FFP", "FFF"):
        mb.trans_many(group, sp)
        for s1 in single:
            mb.trans(s1, sp)
        for q, (g1, s1) in enumerate(zip(group, single)):
            assert g1.space == sp and list(g1.loc_sz) == list(s1.loc_sz)
            same = np.array_equal(g1.download(), s1.download())
            assert same, f"trans_many -> {sp}: scalar {q} differs from trans()"
    if rank == 0:
        print("  trans_many == trans (bit-identical) through PPP/FFF/PFP/PPP/FFP/FFF", flush=True)

    # ---- operators with cross-rank reductions ----
    s = mb.Scalar("FFF").upload_global(e0)
    so = mo.Scalar(e=e0.copy(order="F"), space="FFF")
    g = mb.svv_filter(s, 0.3)
    go = mo.svv_filter(so, ok, 0.3)
    assert abs(g - go) <= 1e-13 * max(1.0, abs(go)), (g, go)
    check("svv_filter", s, so.e)
    c1, c1o = mb.calcat1(s), mo.calcat1(so, ok)
    assert rel_l2(c1, c1o) < TOL
    c0, c0o = mb.calcat0(s), mo.calcat0(so, ok)
    assert rel_l2(c0, c0o) < TOL
    mb.zeroat1(s)
    mo.zeroat1(so, ok)
    check("zeroat1", s, so.e)
    mb.del2(s)
    mo.del2(so, ok)
    check("del2", s, so.e)

    # per-operator parity: every operator starts from the oracle's state, so the 1e-12 bar is not diluted by the
    # conditioning of the operators applied before it
    def resync():
        s.upload_global(so.e)
        s.ln = so.ln

    # ln is one element of the (m = 0, k = 0) column (times exp(lognorm)), broadcast to every rank: it is as sharp as
    # that column -- 1e-12 of the column's norm, and for the Poisson solve eps * cond of the oracle's own matrix
    def ln_close(got, want, col, cond=1.0):
        scale = max(1.0, abs(want), float(np.linalg.norm(col)) * float(np.exp(ok.lognorm[0, 0])))
        return abs(got - want) <= max(TOL, np.finfo(float).eps * cond) * scale

    resync()
    band = mo.leg_del2(0, 0.0, ok.p.nrchop, ok)
    full = np.zeros((ok.p.nrchop, ok.p.nrchop))
    for d, v in band.items():
        ii = np.arange(ok.p.nrchop)
        sel = (ii + d >= 0) & (ii + d < ok.p.nrchop)
        full[ii[sel], ii[sel] + d] = v[sel]
    mb.idel2(s)
    mo.idel2_proln(so, ok)
    check("idel2", s, so.e)
    assert ln_close(s.ln, so.ln, so.e[:, 0, 0], float(np.linalg.cond(full))), (s.ln, so.ln)
    resync()
    mb.idelsqp(s)
    mo.idelsqp(so, ok)
    check("idelsqp", s, so.e)
    assert ln_close(s.ln, so.ln, so.e[:, 0, 0]), (s.ln, so.ln)
    if hp:
        resync()
        mb.ihelmp(s, hp, -3.0e6, 0.5)
        mo.ihelmp(so, hp, -3.0e6, 0.5, ok)
        check("ihelmp", s, so.e)

    # ---- q-vortex: bootstrap + ABCN steps, parity per step ----
    if steps:
        dt = 1e-2
        psi, chi = mv.qvort_dist_tp(kit)
        uz = mv.uniform_z_fld(kit)
        st = mv.bootstrap(kit, dt, psi, chi, uz)
        opsi, ochi = mo.qvort_dist_tp(ok)
        ouz = mo.uniform_z_fld(ok)
        ost = mo.vortex_bootstrap(ok, dt, opsi, ochi, ouz)
        check("bootstrap psi", st.psi, ost.psi.e)
        check("bootstrap chi", st.chi, ost.chi.e)
        for it in range(steps):
            mv.step(st, dt)
            mo.vortex_step(ost, ok, dt)
            check(f"step {it + 1} psi", st.psi, ost.psi.e)
            check(f"step {it + 1} chi", st.chi, ost.chi.e)
    mb.device_sync()
    mb.dist.detach()
    mb.finalize()


def run_qvortex_snapshots(path, rank, world):
    """BASELINE.json configs[2] physics at 64^3 (hyperpow 8, SVV on): Richardson bootstrap + ABCN steps against the
    oracle snapshots tests/oracle_vortex.py wrote (computed once by the test harness, not once per rank)."""
    from oracle_vortex import DT, qvortex_params
    snaps = np.load(path)
    nsnap = snaps["meta"].shape[0]
    n = snaps["psi_0"].shape[0] - 8
    kit = mb.TfmKit.init(qvortex_params(n), rank, world)
    mb.dist.attach()
    if rank == 0:
        print(f"q-vortex {n}^3 hyperpow=8 SVV on, {world} ranks, bootstrap + {nsnap - 1} ABCN steps", flush=True)
    psi, chi = mv.qvort_dist_tp(kit)
    uz = mv.uniform_z_fld(kit)
    check("initial psi", psi, snaps["psi_ic"])
    check("initial chi", chi, snaps["chi_ic"])
    st = mv.bootstrap(kit, DT, psi, chi, uz)
    for it in range(nsnap):
        check(f"{'bootstrap' if it == 0 else f'step {it}'} psi", st.psi, snaps[f"psi_{it}"])
        check(f"{'bootstrap' if it == 0 else f'step {it}'} chi", st.chi, snaps[f"chi_{it}"])
        ln_p, ln_c, g_p, g_c = snaps["meta"][it]
        assert abs(st.psi.ln - ln_p) <= TOL * max(1.0, abs(ln_p)) and abs(st.chi.ln - ln_c) <= TOL * max(1.0, abs(ln_c))
        assert abs(st.gain_psi - g_p) <= TOL and abs(st.gain_chi - g_c) <= TOL, (st.gain_psi, g_p, st.gain_chi, g_c)
        if it + 1 < nsnap:
            mv.step(st, DT)
    mb.device_sync()
    mb.dist.detach()
    mb.finalize()


def run_late_rank_and_io(rank, world, tmpdir):
    """A rank that is seconds late to an exchange must not change any result (the exchange barrier waits),
    and msave/mload of a global file through the C ABI (rank 0 lays the file out, host barrier, everyone fills in its
    slab; submodules/mlegs_scalar_io.f90:6-115) round-trips bit-exactly on several ranks."""
    import time
    p = mb.make_params(32, 16, 8, 32, 9, 5, ell=4.0, zlen=2 * np.pi, visc=1e-4, hyperpow=0, hypervisc=0.0)
    kit = mb.TfmKit.init(p, rank, world)
    mb.dist.attach()
    ok = oracle_kit(kit)
    e0 = random_fff(ok, seed=9)
    s = mb.Scalar("FFF").upload_global(e0)
    so = mo.Scalar(e=e0.copy(order="F"), space="FFF")
    mb.device_sync()
    if rank == world - 1:
        time.sleep(2.5)                      # longer than the 1.1 s the round-1 barrier waited before giving up
    mb.trans(s, "PPP")
    mo.trans(so, "PPP", ok)
    check("trans -> PPP with the last rank 2.5 s late", s, so.e)
    if rank == 0:
        time.sleep(2.5)
    g = mb.svv_filter(s_fff := mb.Scalar("FFF").upload_global(e0), 0.3)
    go = mo.svv_filter(mo.Scalar(e=e0.copy(order="F"), space="FFF"), ok, 0.3)
    assert abs(g - go) <= 1e-13 * max(1.0, abs(go)), (g, go)
    # global file through mlegs_b200_msave / mlegs_b200_mload on every rank; a stale file of the same name must not
    # survive or swallow anybody's slab
    fn = os.path.join(tmpdir, "dist_field.bin")
    if rank == 0:
        with open(fn, "wb") as fh:
            fh.write(b"stale" * 1000)
    mb.dist.allreduce([0.0])
    s_fff.ln = 0.625
    s_fff.chop_offset(1, 0, 2)
    want = s_fff.download()
    for binary in (True, False):
        mb.msave(s_fff, fn, is_binary=binary, is_global=True)
        t = mb.Scalar("FFF")
        mb.mload(fn, t, is_binary=binary, is_global=True)
        assert t.space == "FFF" and t.ln == 0.625 and (t.f.nrchop_offset, t.f.nzchop_offset) == (1, 2)
        if binary:
            assert np.array_equal(t.download(), want), "msave/mload (binary, global) on several ranks"
        else:
            assert rel_l2(t.download(), want) < 1e-15
        mb.dist.allreduce([0.0])             # nobody overwrites the file while another rank still reads it
    if rank == 0:
        print("  late rank + multi-rank msave/mload through the C ABI: ok", flush=True)
    mb.device_sync()
    mb.dist.detach()
    mb.finalize()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    try:
        run_case(32, 16, 8, 32, 9, 5, 4.0, 8, rank, world, steps=2)        # validate_tutorials.py 3-D gate sizes
        run_case(36, 30, 20, 30, 12, 9, 2.0, 4, rank, world, steps=0)      # radix 3/5 lengths, generic FFT kernels
        run_case(64, 64, 64, 64, 33, 33, 4.0, 0, rank, world, steps=0)     # register FFT kernels + compact axial FFT
        run_late_rank_and_io(rank, world, os.environ.get("MLEGS_TEST_TMP", "/tmp"))
        if os.environ.get("MLEGS_QVORTEX_SNAPSHOTS"):
            run_qvortex_snapshots(os.environ["MLEGS_QVORTEX_SNAPSHOTS"], rank, world)
        if rank == 0:
            print("DIST WORKER OK", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
