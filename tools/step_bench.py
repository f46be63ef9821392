#!/usr/bin/env python
"""Time per q-vortex time step (BASELINE.json configs[2]/[3]: input.params physics, ABCN after the Richardson
bootstrap, de-aliasing + SVV) on one GPU or slab-distributed under torchrun, with the per-kernel CUDA-event
breakdown.  Prints one JSON line.

    python tools/step_bench.py --size 256 --steps 5
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/step_bench.py --size 256
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--ncu-step", action="store_true",
                help="bracket ONE extra step with cudaProfilerStart/Stop (ncu --profile-from-start off)")
args = ap.parse_args()

import torch  # noqa: E402
import mlegs_b200 as mb  # noqa: E402
from mlegs_b200 import vortex  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist = None
if world > 1:
    os.environ.pop("NCCL_DEBUG", None)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

n = args.size
# input.params:7-29 physics with NR=NP=NZ=n (SURVEY.md section 8d input 3/4), ran_noise = 0
p = mb.make_params(n, n, n, n, n // 2 + 1, n // 2 + 1, ell=4.0, zlen=2 * np.pi, visc=1e-4, hyperpow=8, hypervisc=5e-7,
                   is_svv=True, svv_cutoff=0.75, svv_target=2e-2, svv_strength=0.12, svv_relax=0.25)
kit = mb.TfmKit.init(p, rank, world)
if world > 1:
    mb.dist.attach()
stream = torch.cuda.Stream()
mb.set_stream(stream.cuda_stream)
dt = 1e-2
psi, chi = vortex.qvort_dist_tp(kit)
uz = vortex.uniform_z_fld(kit)
st = vortex.bootstrap(kit, dt, psi, chi, uz)
for _ in range(args.warmup):
    vortex.step(st, dt)
mb.device_sync()
if dist is not None:
    dist.barrier()
mb.launch_count(reset=True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(stream)
for _ in range(args.steps):
    vortex.step(st, dt)
ev1.record(stream)
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / args.steps
launches = mb.launch_count() / args.steps
if dist is not None:
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
if args.ncu_step:
    mb.device_sync()
    torch.cuda.cudart().cudaProfilerStart()
    vortex.step(st, dt, check=False)
    mb.device_sync()
    torch.cuda.cudart().cudaProfilerStop()
mb.prof_enable(True)
vortex.step(st, dt)
prof = mb.prof_report()
mb.prof_enable(False)
tot = sum(v["ms"] for v in prof.values())
if rank == 0:
    print(json.dumps({"workload": f"q-vortex ABCN step {n}^3 hyperpow=8 SVV on", "n_gpus": world, "ms_per_step": ms,
                      "gdof_steps_per_s": n ** 3 / (ms * 1e-3) / 1e9, "launches_per_step": launches,
                      "finite": bool(mb.is_finite(st.psi)),
                      "kernel_ms_sum": tot,
                      "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / tot, 3)}
                                  for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}), flush=True)
if dist is not None:
    mb.device_sync()
    dist.barrier()
    mb.dist.detach()
    dist.destroy_process_group()
