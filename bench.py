#!/usr/bin/env python
"""bench.py -- transform round-trip throughput (GDOF/s) of the MLegS hot path on B200.

Workload (BASELINE.json configs[1]): 3-D scalar PPP<->FFF transform round trip at NR=NP=NZ=128
(NRCHOP=128, NPCHOP=NZCHOP=65, L=4, ZLEN=2*pi; SURVEY.md section 8d input 2).  A "step" is one
forward + one backward transform of every field of a batch of NF distinct fields; NF is chosen so
the batch (NF x 17.4 MB) is larger than the 126 MB L2, i.e. consecutive kernels never find their
input in cache ("inputs larger than L2").  GDOF/s = NF*NR*NP*NZ / t_step / 1e9.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--size S]

N > 1 is launched by torchrun, one rank per GPU; every field is slab-distributed over all ranks like the
reference's MPI run (one all-to-all per one-way transform, over NVLink peer memory) and the batch grows with
N so that per-GPU work is fixed (weak scaling: NF field-equivalents per GPU).
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


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference path (bench's cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_roundtrip_rate(kit_tables, params, nfields: int, nthreads: int, reps: int):
    """Round trips per second of the NumPy oracle on `nthreads` host threads (one field per thread,
    BLAS pinned to one thread each so threads are the only parallelism)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import mlegs_oracle as mo
    from helpers import random_fff
    try:
        from threadpoolctl import threadpool_limits
    except Exception:   # pragma: no cover
        threadpool_limits = None
    okit = mo.kit_init(params, tables=kit_tables)
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
    with (threadpool_limits(limits=1) if threadpool_limits else contextlib.nullcontext()):
        run()   # warm-up (FFT plans, page faults)
        ts = [run() for _ in range(reps)]
    t = min(ts)
    return nfields / t, t


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Fortran/MPI reference cannot be
    compiled in this image) on all host cores, same workload/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mlegs_b200 as mb
    from helpers import oracle_params
    wl = workload(args.size, args.gpus, args.weak)
    p = mb.make_params(wl["nr"], wl["np"], wl["nz"], wl["nrchop"], wl["npchop"], wl["nzchop"], ell=wl["ell"],
                       zlen=wl["zlen"])
    kit = mb.TfmKit.build_tables(p)          # host-only table build (no GPU work)
    cores = os.cpu_count() or 1
    nthreads = min(cores, 64)
    nfields = nthreads
    dof = wl["nr"] * wl["np"] * wl["nz"]
    # warm-up + timed steps, each a bounded sample: one round trip per thread
    per_step = []
    for i in range(args.warmup + args.steps):
        rate, t = cpu_roundtrip_rate(kit.tables(), oracle_params(p), nfields, nthreads, reps=1)
        if i >= args.warmup:
            per_step.append((rate, t))
    rate = float(np.mean([r for r, _ in per_step]))
    val = rate * dof / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean([t for _, t in per_step]) * 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"PPP<->FFF round trip {wl['nr']}x{wl['np']}x{wl['nz']}",
                       "fields_per_step": nfields, **{k: (float(v) if isinstance(v, float) else v) for k, v in wl.items()}},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port",
                             "sample": f"{nfields} fields x 1 round trip per step, NumPy oracle port "
                                       "(pocketfft + BLAS), one field per thread"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import mlegs_b200 as mb
    from helpers import oracle_kit, oracle_params, random_fff
    from oracle import mlegs_oracle as mo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)     # its "NCCL version" banner goes to stdout; stdout carries ONE JSON line
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
    field_bytes = int(np.prod(kit.glb_sz)) * 16
    # fields per GPU-equivalent: the batch is > 2x L2 and long enough for the clock sampler to see it
    gpu_bytes = field_bytes // world      # one field's share on one GPU
    nf = args.fields or max(2, int(np.ceil(2.2 * 126e6 / gpu_bytes)), min(64, int(1.2e9 // gpu_bytes)))
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
        sampler.rows.clear()      # keep only the samples of the timed region
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
    clocks = sampler.stop() if rank == 0 else None
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
    hbm_peak, hbm_src = measured_peaks()
    dmma_peak = mb.dmma_peak()
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
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = kernels[dom]
    if dom.startswith("legendre") and d["fp64_tensor_frac"] >= d["hbm_frac"]:
        roofline = {"kernel": dom, "bound": "tensor", "achieved": d["TFLOPs"], "peak": dmma_peak, "unit": "TFLOP/s",
                    "frac": d["fp64_tensor_frac"], "traffic": None,
                    "peak_source": "FP64 DMMA m8n8k4 micro-benchmark measured live in this run "
                                   "(mlegs_b200_dmma_peak; MEASURED_PEAKS.json has no FP64 entry)"}
    else:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": d["alg_GBps"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": d["hbm_frac"], "traffic": None, "peak_source": hbm_src}
    roofline.update({"avg_launch_ms": d["avg_ms"], "share_of_step": d["share"], "dmma_peak_tflops": dmma_peak,
                     "kernels": kernels})
    traffic_file = os.path.join(ROOT, "profiles", "r1", "traffic.json")
    if os.path.exists(traffic_file) and world == 1 and SHAPE is None:
        tr = json.load(open(traffic_file)).get(str(args.size), {})
        per_field = tr.get(dom)
        roofline["traffic"] = per_field * d["scalars_per_launch"] if per_field is not None else None
        roofline["traffic_source"] = (str(tr.get("source")) + "; per-scalar figure x scalars_per_launch")

    # ---- e2e: the reference-facing host-buffer entry, pinned host arrays, H2D+D2H inside the timed region ----
    n_ppp = int(np.prod(fields[0].loc_sz))
    mb.trans(fields[0], "FFF")
    n_fff = int(np.prod(fields[0].loc_sz))
    mb.trans(fields[0], "PPP")
    nhost = max(n_ppp, n_fff)
    ne2e = min(len(fields), 16 * world)
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
    e2e = {"value": ne2e * dof / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": ne2e * 16 * (n_ppp + n_fff), "d2h_bytes_per_step": ne2e * 16 * (n_ppp + n_fff),
           "ms_per_step": e2e_ms, "fields_per_step": ne2e,
           "api": "mlegs_b200_trans_host_batch (host s%e in, host s%e out for every field of the batch; H2D, "
                  "transform and D2H pipelined), pinned host arrays; bytes are per rank"}

    # ---- CPU baseline on rank 0: bounded sample of the same workload with the oracle port ----
    cpu = None
    if rank == 0 and not args.no_cpu:
        nthreads = min(os.cpu_count() or 1, 16)
        rate, t = cpu_roundtrip_rate(kit.tables(), oracle_params(p), nthreads, nthreads, reps=2)
        cpu = {"value": rate * dof / 1e9, "unit": UNIT, "cores": nthreads, "kind": "port",
               "sample": f"{nthreads} fields x 1 round trip (best of 2), NumPy oracle port of ops:157-235, "
                         "one field per thread"}

    if rank == 0:
        par = "single GPU" if world == 1 else (f"every field slab-distributed over {world} GPUs (r / m shards), "
                                               "one peer-memory all-to-all per one-way transform")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": (f"PPP<->FFF round trip {wl['nr']}x{wl['np']}x{wl['nz']} "
                                        + ("(BASELINE.json configs[1])" if world == 1 else
                                           f"(configs[1] extended along the periodic axis: {args.size}^3 DOF per GPU, "
                                           "BASELINE.json configs[4] sweep shape)" if args.weak == "nz" else
                                           "(BASELINE.json configs[1], batch grown with N)")),
                           "fields_per_step": nfields, "scalars_per_launch": nb,
                           "l2_policy": f"inputs larger than L2: {nfields} distinct fields x {field_bytes / 1e6:.1f} MB "
                                        f"({nfields * gpu_bytes / 1e6:.0f} MB per GPU)",
                           "parallelism": par,
                           **{k: (float(v) if isinstance(v, float) else v) for k, v in wl.items()}},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e}
        print(json.dumps(line), flush=True)
    if dist is not None:
        mb.device_sync()
        dist.barrier()
        mb.dist.detach()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--fields", type=int, default=0)
    ap.add_argument("--batch", type=int, default=8,
                    help="scalars per mlegs_b200_trans_many call (1: one mlegs_b200_trans per scalar)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--shape", default="", help="NR,NP,NZ of a non-cubic sweep point (overrides --size)")
    ap.add_argument("--weak", default="nz", choices=["nz", "fields"],
                    help="N > 1: grow NZ with N (default, DOF per GPU fixed) or grow the batch of cubic fields")
    args = ap.parse_args()
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
