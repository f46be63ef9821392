#!/usr/bin/env python
"""bench.py -- the two halves of BASELINE.json's metric on B200: transform round-trip GDOF/s and q-vortex time/step.

Headline (`value`, BASELINE.json configs[1]): 3-D scalar PPP<->FFF transform round trip at NR=NP=NZ=128
(NRCHOP=128, NPCHOP=NZCHOP=65, L=4, ZLEN=2*pi; SURVEY.md section 8d input 2).  A "step" is one forward + one backward
transform of every field of a batch of NF distinct fields; NF is chosen so the batch (NF x 17.4 MB) is larger than the
126 MB L2, i.e. consecutive kernels never find their input in cache ("inputs larger than L2").
GDOF/s = NF*NR*NP*NZ / t_step / 1e9.  N > 1 (torchrun, one rank per GPU): every field is slab-distributed over all
ranks like the reference's MPI run (one all-to-all per one-way transform, over NVLink peer memory) and the axial
direction grows with N so that per-GPU work is fixed (weak scaling).

`time_step` (BASELINE.json configs[2]): the q-vortex ABCN step of apps/vortical_flow_3d.f90:160-206 with the
input.params physics (hyperpow 8, SVV on, de-aliasing, ran_noise = 0) at NR=NP=NZ=256, on the same N GPUs (strong
scaling: the problem is fixed, every field is slab-distributed), ms/step with CUDA events after the Richardson bootstrap,
plus the per-kernel roofline of one step.  At N = 8 `time_step_512` adds configs[3] (512^3, the north-star target).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--size S]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "transform_roundtrip_gdofs"
UNIT = "GDOF/s"
# measured peer-copy bandwidth per direction per GPU on this pool (B200_PROFILING.md; nominal 900 GB/s)
NVLINK_PEER_GBPS = 770.0
# FP64 tensor pipe: DMMA m8n8k4 issues 64 FMA per clock per SM (tools/microbench/dmma_shapes.cu measures 63.8)
DMMA_FMA_PER_CLK_SM = 64
N_SMS = 148

SHAPE = None   # --shape NR,NP,NZ (BASELINE.json configs[4]: the sweep's non-cubic points)


def workload(size: int, world: int = 1, weak: str = "nz"):
    """N = 1: the cubic case.  N > 1, weak scaling: the periodic axial direction is extended with the GPU count
    (NZ = size * N, ZLEN scaled alike), so every GPU keeps size^3 degrees of freedom of every field and the same
    Legendre/FFT work per field as the single-GPU run; --weak fields keeps the cube and grows the batch instead."""
    if SHAPE is not None:
        nr, npp, nz0 = SHAPE
        nz = nz0 * world if weak == "nz" else nz0
        return dict(nr=nr, np=npp, nz=nz, nrchop=nr, npchop=npp // 2 + 1, nzchop=nz // 2 + 1,
                    ell=4.0, zlen=2.0 * np.pi * max(1, nz // nz0))
    nz = size * world if weak == "nz" else size
    return dict(nr=size, np=size, nz=nz, nrchop=size, npchop=size // 2 + 1, nzchop=nz // 2 + 1,
                ell=4.0, zlen=2.0 * np.pi * (nz // size))


def fields_per_step(wl, world: int, override: int = 0):
    """Batch size of one step: > 2x L2 per GPU and long enough for the clock sampler to see it."""
    nrdim = wl["nr"] + 3
    field_bytes = nrdim * (wl["np"] // 2 + 1) * wl["nz"] * 16
    gpu_bytes = field_bytes // world
    nf = override or max(2, int(np.ceil(2.2 * 126e6 / gpu_bytes)), min(64, int(1.2e9 // gpu_bytes)))
    return nf, field_bytes, gpu_bytes


def bench_config(args, wl, world: int, nfields: int, field_bytes: int, gpu_bytes: int):
    """The `config` object: identical for the native and the reference arm of one command line."""
    if world == 1:
        what = "(BASELINE.json configs[1])"
        par = "single GPU"
    else:
        what = (f"(configs[1] extended along the periodic axis: {args.size}^3 DOF per GPU, BASELINE.json configs[4] "
                "sweep shape)" if args.weak == "nz" else "(BASELINE.json configs[1], batch grown with N)")
        par = (f"every field slab-distributed over {world} GPUs (r / m shards), one peer-memory all-to-all per one-way "
               "transform")
    return {"workload": f"PPP<->FFF round trip {wl['nr']}x{wl['np']}x{wl['nz']} {what}",
            "fields_per_step": nfields, "scalars_per_launch": max(1, args.batch),
            "l2_policy": f"inputs larger than L2: {nfields} distinct fields x {field_bytes / 1e6:.1f} MB "
                         f"({nfields * gpu_bytes / 1e6:.0f} MB per GPU)",
            "parallelism": par, **{k: (float(v) if isinstance(v, float) else v) for k, v in wl.items()}}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.rows = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def summary(self, rows):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            parts = [x.strip() for x in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}

    def mark(self):
        """Samples collected so far are dropped; returns nothing."""
        self.rows.clear()

    def snapshot(self):
        time.sleep(0.05)
        return self.summary(list(self.rows))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        return self.summary(self.rows)


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference path (bench's cpu_baseline leg and --impl reference).
# Nothing in this section imports mlegs_b200.
# ------------------------------------------------------------------------------------------------
def oracle_kit_cached(wl, workers: int):
    """Oracle kit with its own tables (oracle.kit_init: GL nodes, lognorm, the 50-digit recurrence of sinit:254-300),
    cached on disk under oracle/_cache keyed by (nr, nrchop, npchop) -- the table does not depend on nz, ell or zlen."""
    from oracle import mlegs_oracle as mo
    p = mo.Params(nr=wl["nr"], np=wl["np"], nz=wl["nz"], nrchop=wl["nrchop"], npchop=wl["npchop"], nzchop=wl["nzchop"],
                  ell=wl["ell"], zlen=wl["zlen"])
    cdir = os.path.join(ROOT, "oracle", "_cache")
    path = os.path.join(cdir, f"tables_{p.nr}_{p.nrchop}_{p.npchop}.npz")
    if os.path.exists(path):
        t = np.load(path)
        return mo.kit_init(p, tables={k: t[k] for k in ("x", "w", "lognorm", "pf", "at0", "at1")}), p
    mo.set_workers(workers)
    try:
        kit = mo.kit_init(p)
    finally:
        mo.set_workers(1)
    try:
        os.makedirs(cdir, exist_ok=True)
        np.savez(path + ".tmp.npz", x=kit.x, w=kit.w, lognorm=kit.lognorm, pf=kit.pf, at0=kit.at0, at1=kit.at1)
        os.replace(path + ".tmp.npz", path)
    except OSError:
        pass
    return kit, p


def cpu_roundtrip_rate(okit, nfields: int, nthreads: int, reps: int, check_ppp=None):
    """Round trips per second of the NumPy oracle on `nthreads` host threads (one field per thread, BLAS pinned to one
    thread each so threads are the only parallelism).  check_ppp: a PPP field whose oracle forward transform is
    returned as well (the checker of the native arm's `parity` entry)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import mlegs_oracle as mo
    from helpers import random_fff
    try:
        from threadpoolctl import threadpool_limits
    except Exception:   # pragma: no cover
        threadpool_limits = None
    base = mo.Scalar(e=random_fff(okit, seed=0), space="FFF")
    mo.trans(base, "PPP", okit)
    fields = [base.copy() for _ in range(nfields)]

    def work(s):
        mo.trans(s, "FFF", okit)
        mo.trans(s, "PPP", okit)

    def run():
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=nthreads) as ex:
            list(ex.map(work, fields))
        return time.perf_counter() - t0

    import contextlib
    checked = None
    with (threadpool_limits(limits=1) if threadpool_limits else contextlib.nullcontext()):
        run()   # warm-up (FFT plans, page faults)
        ts = [run() for _ in range(reps)]
        if check_ppp is not None:
            so = mo.Scalar(e=np.asfortranarray(check_ppp), space="PPP")
            mo.trans(so, "FFF", okit)
            checked = so.e
    t = min(ts)
    return nfields / t, t, checked


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Fortran/MPI reference cannot be compiled in
    this image) on all host cores, same workload/metric/config as the native arm.  Pure oracle: no mlegs_b200 import."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.size, args.gpus, args.weak)
    cores = os.cpu_count() or 1
    nthreads = min(cores, 64)
    okit, _ = oracle_kit_cached(wl, nthreads)
    nfields, field_bytes, gpu_bytes = fields_per_step(wl, args.gpus, args.fields)
    if args.weak != "nz":
        nfields *= args.gpus
    # bounded sample of the step: the whole batch at N = 1 (about a second per step), one field per thread beyond
    sample = nfields if args.gpus == 1 else min(nfields, nthreads)
    dof = wl["nr"] * wl["np"] * wl["nz"]
    per_step = []
    for i in range(args.warmup + args.steps):
        rate, t, _ = cpu_roundtrip_rate(okit, sample, nthreads, reps=1)
        if i >= args.warmup:
            per_step.append((rate, t))
    rate = float(np.mean([r for r, _ in per_step]))
    val = rate * dof / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean([t for _, t in per_step]) * 1e3 * nfields / sample),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, wl, args.gpus, nfields, field_bytes, gpu_bytes),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port",
                             "sample": f"{sample} of the step's {nfields} fields x 1 round trip per step, NumPy oracle "
                                       "port (pocketfft + BLAS), one field per thread; tables from oracle.kit_init; "
                                       "ms_per_step is scaled to the whole batch"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def kernel_table(prof: dict, hbm_peak: float, dmma_peak: float, world: int = 1, nvlink_bytes=None):
    """Per-kernel roofline entries from the library's CUDA-event profile (ms, algorithmic bytes and flops per kernel)."""
    tot = sum(v["ms"] for v in prof.values()) or 1.0
    out = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        ms = v["ms"]
        ent = {"launches": v["launches"], "ms": round(ms, 4), "share": round(ms / tot, 4)}
        if v.get("bytes", 0) > 0:
            ent["alg_GBps"] = round(v["bytes"] / (ms * 1e-3) / 1e9, 1)
            ent["hbm_frac"] = round(ent["alg_GBps"] / hbm_peak, 3)
        if v.get("flops", 0) > 0:
            ent["TFLOPs"] = round(v["flops"] / (ms * 1e-3) / 1e12, 2)
            ent["fp64_tensor_frac"] = round(ent["TFLOPs"] / dmma_peak, 3)
        if nvlink_bytes and k in nvlink_bytes:
            ent["nvlink_GBps"] = round(nvlink_bytes[k] / (ms * 1e-3) / 1e9, 1)
            ent["nvlink_frac"] = round(ent["nvlink_GBps"] / NVLINK_PEER_GBPS, 3)
        out[k] = ent
    return out


def run_time_step(mb, torch, dist, size: int, rank: int, world: int, stream, steps: int, warmup: int, sampler,
                  hbm_peak: float, dmma_peak: float):
    """q-vortex ABCN step (apps/vortical_flow_3d.f90:160-206) at size^3 with the input.params physics on `world` GPUs."""
    from mlegs_b200 import vortex
    n = size
    p = mb.make_params(n, n, n, n, n // 2 + 1, n // 2 + 1, ell=4.0, zlen=2 * np.pi, visc=1e-4, hyperpow=8,
                       hypervisc=5e-7, is_svv=True, svv_cutoff=0.75, svv_target=2e-2, svv_strength=0.12, svv_relax=0.25)
    t_init = time.perf_counter()
    kit = mb.TfmKit.init(p, rank, world)
    if world > 1:
        mb.dist.attach()
    mb.set_stream(stream.cuda_stream)
    dt = 1e-2
    psi, chi = vortex.qvort_dist_tp(kit)
    uz = vortex.uniform_z_fld(kit)
    st = vortex.bootstrap(kit, dt, psi, chi, uz)
    for _ in range(max(warmup, 3)):
        vortex.step(st, dt)
    mb.device_sync()
    t_init = time.perf_counter() - t_init
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.mark()
    mb.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        vortex.step(st, dt)
    ev1.record(stream)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    launches = mb.launch_count() / steps
    clocks = sampler.snapshot() if sampler is not None else None
    finite = bool(mb.is_finite(st.psi) and mb.is_finite(st.chi))
    if dist is not None:
        t = torch.tensor([ms, 0.0 if finite else 1.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, finite = float(t[0].item()), bool(t[1].item() == 0.0)
    mb.prof_enable(True)
    vortex.step(st, dt, check=False)
    prof = mb.prof_report()
    mb.prof_enable(False)
    kernels = kernel_table(prof, hbm_peak, dmma_peak, world)
    out = {"workload": f"q-vortex ABCN step {n}x{n}x{n} (BASELINE.json configs[{2 if n == 256 else 3}]: input.params "
                       "physics, hyperpow 8, SVV on, de-aliasing, ran_noise 0; apps/vortical_flow_3d.f90:160-206)",
           "n_gpus": world, "scaling": "strong", "ms_per_step": ms, "steps": steps,
           "gdof_steps_per_s": n ** 3 / (ms * 1e-3) / 1e9, "gpu_launches_per_step": launches, "finite": finite,
           "kernel_ms_sum": round(sum(v["ms"] for v in prof.values()), 4), "setup_s": round(t_init, 2),
           "clocks": clocks, "kernels": kernels,
           "parity": "bootstrap + 3 steps of this configuration are checked against the oracle at 128^3 on one GPU "
                     "(tests/test_gpu_step_parity.py) and at 64^3 on 2/4/8 GPUs (tests/test_gpu_dist.py), 1e-12 per step"}
    # free the time-step state before the next kit
    del st, psi, chi, uz
    mb.device_sync()
    if dist is not None:
        dist.barrier()
    if world > 1:
        mb.dist.detach()
    mb.finalize()
    return out


def run_native(args):
    import torch
    import mlegs_b200 as mb
    from helpers import oracle_kit, random_fff, rel_l2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # stdout carries ONE JSON line: NCCL's own log (NCCL_DEBUG=INFO prints to stdout by default) goes to stderr
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = workload(args.size, world, args.weak)
    p = mb.make_params(wl["nr"], wl["np"], wl["nz"], wl["nrchop"], wl["npchop"], wl["nzchop"], ell=wl["ell"],
                       zlen=wl["zlen"])
    # N > 1: every field is slab-distributed over all ranks like the reference's MPI run (PPP sharded in r,
    # spectral spaces in m); each one-way transform then contains one all-to-all over NVLink peer memory.
    kit = mb.TfmKit.init(p, rank, world)
    if world > 1:
        mb.dist.attach()
    dof = wl["nr"] * wl["np"] * wl["nz"]
    nf, field_bytes, gpu_bytes = fields_per_step(wl, world, args.fields)
    assert field_bytes == int(np.prod(kit.glb_sz)) * 16
    # weak scaling, per-GPU work fixed: either the fields grow with N (default) or the batch does
    nfields = nf if args.weak == "nz" else nf * world
    stream = torch.cuda.Stream()
    mb.set_stream(stream.cuda_stream)

    okit = oracle_kit(kit)
    e0 = random_fff(okit, seed=0)
    fields = []
    with torch.cuda.stream(stream):
        s0 = mb.Scalar("FFF").upload_global(e0)
        mb.trans(s0, "PPP")
        fields.append(s0)
        for _ in range(nfields - 1):
            fields.append(s0.copy())
    mb.device_sync()
    # parity record of this very run: field 0 before the timed loop, forward-transformed by the device (below) and
    # by the oracle (inside the cpu_baseline leg)
    check_ppp = check_fff = None
    if world == 1:
        check_ppp = s0.download()
        c0 = s0.copy()
        mb.trans(c0, "FFF")
        check_fff = c0.download()
        del c0

    # mlegs_b200_trans_many runs every stage of a group of scalars as one launch (scalar index = a grid dimension);
    # on several ranks the group also shares ONE fused peer-memory exchange (and one barrier) per one-way transform
    nb = max(1, args.batch)
    groups = [fields[i:i + nb] for i in range(0, len(fields), nb)]

    def step():
        if nb == 1:
            for s in fields:
                mb.trans(s, "FFF")
                mb.trans(s, "PPP")
        else:
            for g in groups:
                mb.trans_many(g, "FFF")
                mb.trans_many(g, "PPP")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if rank == 0:
        sampler.mark()            # keep only the samples of the timed region
    mb.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = mb.launch_count()
    clocks = sampler.snapshot() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = nfields * dof / (ms_step * 1e-3) / 1e9

    # ---- per-kernel CUDA-event timing of the same workload (separate pass, not the headline number) ----
    mb.prof_enable(True)
    prof_steps = 2
    for _ in range(prof_steps):
        step()
    prof = mb.prof_report()
    mb.prof_enable(False)
    tot = sum(v["ms"] for v in prof.values())
    nr, npn, nz = wl["nr"], wl["np"], wl["nz"]
    nrdim, npdim = kit.glb_sz[0], kit.glb_sz[1]
    S = sum(max(wl["nrchop"] - m, 0) for m in range(wl["npchop"]))
    # algorithmic bytes / flops per launch of ONE whole field; a rank
    # of a slab-distributed run processes 1/world of that per launch
    alg_bytes = {
        "fft_phi_forward": 8 * nr * npn * nz + 16 * nr * npdim * nz,
        "fft_phi_backward": 8 * nr * npn * nz + 16 * nr * npdim * nz,
        "fft_z_forward": 2 * 16 * nz * S,
        "fft_z_backward": 2 * 16 * nz * S,
        "legendre_forward": 16 * nr * wl["npchop"] * nz + 16 * S * nz,
        "legendre_backward": 16 * nr * wl["npchop"] * nz + 16 * S * nz,
        "exchange_21": 2 * 16 * nrdim * npdim * nz,
        "exchange_12": 2 * 16 * nrdim * npdim * nz,
    }
    leg_flops = 2.0 * nr * nz * S
    hbm_peak, hbm_src, peaks = measured_peaks()
    # FP64 tensor peak: pinned to the pipe's issue rate at the maximum SM clock (64 FMA/clk/SM x 148 SMs x 2 flop x
    # sm_max_mhz); the live DMMA micro-benchmark is reported next to it (it reads 37.1 TFLOP/s when the box holds
    # 1965 MHz and less when the probe itself runs into the power limit, which made it useless as a denominator)
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    dmma_peak = 2.0 * DMMA_FMA_PER_CLK_SM * N_SMS * sm_max * 1e6 / 1e12
    dmma_live = mb.dmma_peak()
    kernels = {}
    for k, v in prof.items():
        avg_ms = v["ms"] / v["launches"]
        # scalars one launch of this kernel processes (trans_many: a group per launch)
        per_launch = prof_steps * nfields / v["launches"]
        ent = {"avg_ms": avg_ms, "share": v["ms"] / tot, "launches": v["launches"], "scalars_per_launch": per_launch}
        # *_put: the same kernel with the exchange fused into its stores; *_stage / *_ship: the two halves of the
        # staged (1,2) exchange (rows written locally in destination order, then shipped as long runs)
        base = k.replace("_put", "").replace("_stage", "").replace("_ship", "")
        ent["alg_GBps"] = per_launch * alg_bytes.get(base, 0) / world / (avg_ms * 1e-3) / 1e9
        ent["hbm_frac"] = ent["alg_GBps"] / hbm_peak
        if k.startswith("legendre"):
            ent["TFLOPs"] = per_launch * leg_flops / world / (avg_ms * 1e-3) / 1e12
            ent["fp64_tensor_frac"] = ent["TFLOPs"] / dmma_peak
        if k.endswith("_put") or k.startswith("exchange"):
            # the kernel's stores are the all-to-all: (world-1)/world of this rank's slab of every scalar crosses NVLink
            nv_bytes = per_launch * 16 * nrdim * npdim * nz / world * (world - 1) / world
            ent["nvlink_GBps"] = nv_bytes / (avg_ms * 1e-3) / 1e9
            ent["nvlink_frac"] = ent["nvlink_GBps"] / NVLINK_PEER_GBPS
        kernels[k] = ent
    if dist is not None:
        # load balance over ranks (the m distribution): min / max of every kernel's time
        allk = [None] * world
        dist.all_gather_object(allk, {k: v["ms"] for k, v in prof.items()})
        for k, ent in kernels.items():
            vals = [d.get(k, 0.0) for d in allk]
            ent["ms_min_over_ranks"] = min(vals) / prof[k]["launches"]
            ent["ms_max_over_ranks"] = max(vals) / prof[k]["launches"]
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = kernels[dom]
    if dom.startswith("legendre") and d["fp64_tensor_frac"] >= d["hbm_frac"]:
        roofline = {"kernel": dom, "bound": "tensor", "achieved": d["TFLOPs"], "peak": dmma_peak, "unit": "TFLOP/s",
                    "frac": d["fp64_tensor_frac"], "traffic": None,
                    "peak_source": f"FP64 DMMA issue rate: {DMMA_FMA_PER_CLK_SM} FMA/clk/SM x {N_SMS} SMs x 2 x "
                                   f"{sm_max:.0f} MHz (MEASURED_PEAKS.json and B200_PROFILING.md have no FP64 entry; "
                                   "live micro-benchmark in dmma_live_tflops)"}
    else:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": d["alg_GBps"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": d["hbm_frac"], "traffic": None, "peak_source": hbm_src}
    roofline.update({"avg_launch_ms": d["avg_ms"], "share_of_step": d["share"], "dmma_peak_tflops": dmma_peak,
                     "dmma_live_tflops": dmma_live, "kernels": kernels})
    for rnd in ("r2", "r1"):
        traffic_file = os.path.join(ROOT, "profiles", rnd, "traffic.json")
        if os.path.exists(traffic_file) and world == 1 and SHAPE is None:
            tr = json.load(open(traffic_file)).get(str(args.size), {})
            per_field = tr.get(dom)
            if per_field is not None:
                roofline["traffic"] = per_field * d["scalars_per_launch"]
                roofline["traffic_source"] = (str(tr.get("source")) + "; per-scalar figure x scalars_per_launch")
                break

    # ---- the same step launched 8 scalars at a time: what one exchange epoch carries on several GPUs, so a multi-GPU
    # line can be set against a one-GPU figure of the same launch size (separate pass, not the headline number) ----
    at8 = None
    if world == 1 and nb > 8:
        try:
            groups8 = [fields[i:i + 8] for i in range(0, len(fields), 8)]

            def step8():
                for g in groups8:
                    mb.trans_many(g, "FFF")
                    mb.trans_many(g, "PPP")

            for _ in range(3):
                step8()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(args.steps):
                step8()
            a1.record(stream)
            torch.cuda.synchronize()
            ms8 = a0.elapsed_time(a1) / args.steps
            at8 = {"scalars_per_launch": 8, "ms_per_step": ms8, "value": nfields * dof / (ms8 * 1e-3) / 1e9,
                   "unit": UNIT}
        except Exception as exc:      # never lose the line over the extra figure
            at8 = {"error": repr(exc)}
        finally:
            groups8 = None            # the fields are released with the round-trip state below

    # ---- e2e: the reference-facing host-buffer entry, pinned host arrays, H2D+D2H inside the timed region ----
    n_ppp = int(np.prod(fields[0].loc_sz))
    mb.trans(fields[0], "FFF")
    n_fff = int(np.prod(fields[0].loc_sz))
    mb.trans(fields[0], "PPP")
    nhost = max(n_ppp, n_fff)
    # the same batch as the device-timed step (64 fields at the default size): the three-deep pipeline of the host
    # entry fills and drains once per call, which a 16-field batch paid twice per 13 ms
    ne2e = min(len(fields), 64)
    hosts = []
    for s in fields[:ne2e]:
        h = torch.empty(nhost * 2, dtype=torch.float64).pin_memory()
        a = h.numpy().view(np.complex128)
        a[:n_ppp] = s.download().ravel(order="F")
        hosts.append((h, a))

    harr = [a for _, a in hosts]

    def e2e_step():
        mb.trans_host_batch(harr, "PPP", "FFF")
        mb.trans_host_batch(harr, "FFF", "PPP")

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    # what PCIe gives this GPU with both directions busy (pinned 64 MB buffers, copy engines): the ceiling of the e2e rate
    duplex = None
    if world == 1:
        nb_ = 64 << 20
        hi, ho = torch.empty(nb_, dtype=torch.uint8).pin_memory(), torch.empty(nb_, dtype=torch.uint8).pin_memory()
        di, do = torch.empty(nb_, dtype=torch.uint8, device="cuda"), torch.empty(nb_, dtype=torch.uint8, device="cuda")
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(2 + 8):
            if it == 2:
                torch.cuda.synchronize()
                ev0.record()
                sa.wait_event(ev0)
                sb.wait_event(ev0)
            with torch.cuda.stream(sa):
                di.copy_(hi, non_blocking=True)
            with torch.cuda.stream(sb):
                ho.copy_(do, non_blocking=True)
        torch.cuda.current_stream().wait_stream(sa)
        torch.cuda.current_stream().wait_stream(sb)
        ev1.record()
        torch.cuda.synchronize()
        duplex = nb_ / (ev0.elapsed_time(ev1) / 8 * 1e-3) / 1e9
        del hi, ho, di, do
    e2e_bytes = ne2e * 16 * (n_ppp + n_fff)
    e2e = {"value": ne2e * dof / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": e2e_bytes, "d2h_bytes_per_step": e2e_bytes,
           "ms_per_step": e2e_ms, "fields_per_step": ne2e,
           "pcie_GBps_per_direction": e2e_bytes / (e2e_ms * 1e-3) / 1e9,
           "pcie_duplex_peak_GBps": duplex,
           "pcie_frac": (e2e_bytes / (e2e_ms * 1e-3) / 1e9 / duplex) if duplex else None,
           "api": "mlegs_b200_trans_host_batch (host s%e in, host s%e out for every field of the batch; H2D, "
                  "transform and D2H pipelined), pinned host arrays; bytes are per rank"}

    # ---- CPU baseline on rank 0: bounded sample of the same workload with the oracle port; the same leg checks the
    # device's forward transform of field 0 against the oracle's ----
    cpu = parity = None
    if rank == 0 and not args.no_cpu:
        nthreads = min(os.cpu_count() or 1, 16)
        rate, t, want = cpu_roundtrip_rate(okit, nthreads, nthreads, reps=2, check_ppp=check_ppp)
        cpu = {"value": rate * dof / 1e9, "unit": UNIT, "cores": nthreads, "kind": "port",
               "sample": f"{nthreads} fields x 1 round trip (best of 2), NumPy oracle port of ops:157-235, "
                         "one field per thread"}
        if want is not None:
            parity = {"check": "PPP->FFF of field 0 of this run: device vs oracle, relative L2", "tol": 1e-12,
                      "rel_l2": rel_l2(check_fff, want)}
            parity["ok"] = bool(parity["rel_l2"] < parity["tol"])

    # ---- the round-trip state goes away; the time step runs on a kit of its own ----
    del fields, groups, s0, hosts, harr
    mb.device_sync()
    if dist is not None:
        dist.barrier()
    if world > 1:
        mb.dist.detach()
    mb.finalize()
    tsteps = {}
    if not args.no_step and SHAPE is None:
        k_steps = max(3, min(args.steps, 10))
        tsteps["time_step"] = run_time_step(mb, torch, dist, args.step_size, rank, world, stream, k_steps,
                                            args.warmup, sampler if rank == 0 else None, hbm_peak, dmma_peak)
        if world == 8 and not args.no_step512:
            tsteps["time_step_512"] = run_time_step(mb, torch, dist, 512, rank, world, stream, k_steps, args.warmup,
                                                    sampler if rank == 0 else None, hbm_peak, dmma_peak)
    if rank == 0:
        sampler.stop()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": bench_config(args, wl, world, nfields, field_bytes, gpu_bytes),
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "parity": parity, "value_at_8_per_launch": at8, **tsteps}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--fields", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0,
                    help="scalars per mlegs_b200_trans_many call (1: one mlegs_b200_trans per scalar; 0 = default: 32 on "
                         "one GPU -- a launch costs ~9 us before its first byte moves, tools/rt_bench.py -- and 8 on "
                         "several, what one exchange epoch carries)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-step", action="store_true", help="skip the q-vortex time-step section")
    ap.add_argument("--no-step512", action="store_true", help="N = 8: skip the 512^3 time step")
    ap.add_argument("--step-size", type=int, default=256, help="NR = NP = NZ of the time-step section")
    ap.add_argument("--shape", default="", help="NR,NP,NZ of a non-cubic sweep point (overrides --size)")
    ap.add_argument("--weak", default="nz", choices=["nz", "fields"],
                    help="N > 1: grow NZ with N (default, DOF per GPU fixed) or grow the batch of cubic fields")
    args = ap.parse_args()
    if args.batch <= 0:
        args.batch = 32 if int(os.environ.get("WORLD_SIZE", "1")) == 1 else 8
    if args.shape:
        global SHAPE
        SHAPE = tuple(int(v) for v in args.shape.split(","))
        assert len(SHAPE) == 3, "--shape NR,NP,NZ"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
