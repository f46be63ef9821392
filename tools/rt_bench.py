#!/usr/bin/env python
"""Times the kernels of the PPP <-> FFF round trip on one GPU: groups of `--batch` scalars at size^3 through trans_many,
CUDA events per launch from the library profiler.  For kernel experiments (environment knobs are read at launch).

    python tools/rt_bench.py --size 128 [--fields 64] [--reps 5] [--tag x]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--fields", type=int, default=64)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--tag", default="")
args = ap.parse_args()

import mlegs_b200 as mb  # noqa: E402
from helpers import oracle_kit, random_fff  # noqa: E402

n = args.size
p = mb.make_params(n, n, n, n, n // 2 + 1, n // 2 + 1, ell=4.0, zlen=2 * np.pi)
kit = mb.TfmKit.init(p)
ok = oracle_kit(kit)
s0 = mb.Scalar("FFF").upload(random_fff(ok, seed=0))
ref = s0.download()
fields = [s0] + [s0.copy() for _ in range(args.fields - 1)]
groups = [fields[i:i + args.batch] for i in range(0, len(fields), args.batch)]


def step():
    for g in groups:
        mb.trans_many(g, "PPP")
        mb.trans_many(g, "FFF")


for _ in range(3):
    step()
mb.device_sync()
mb.prof_enable(True)
for _ in range(args.reps):
    step()
prof = mb.prof_report()
mb.prof_enable(False)
out = {"tag": args.tag, "size": n, "batch": args.batch}
tot = 0.0
for k, v in prof.items():
    us = v["ms"] / v["launches"] * 1e3
    tot += us
    out[k] = round(us, 2)
out["sum_us"] = round(tot, 1)
err = np.linalg.norm(fields[-1].download() - ref) / np.linalg.norm(ref)
out["roundtrip_rel_l2"] = float(err)
print(json.dumps(out), flush=True)
