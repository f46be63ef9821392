#!/usr/bin/env python
"""One warm-up + N round trips of a single scalar field (for ncu captures; never a bench value).

    ncu --set full --import-source on -k regex:leg_ -s 2 -c 2 -o gpurun_out/leg512 python tools/prof_roundtrip.py --size 512
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--hyperpow", type=int, default=0)
ap.add_argument("--batch", type=int, default=1, help="scalars per mlegs_b200_trans_many call")
args = ap.parse_args()

import mlegs_b200 as mb  # noqa: E402

n = args.size
p = mb.make_params(n, n, n, n, n // 2 + 1, n // 2 + 1, ell=4.0, zlen=2 * np.pi, hyperpow=args.hyperpow,
                   hypervisc=(5e-7 if args.hyperpow else 0.0))
kit = mb.TfmKit.init(p)
rng = np.random.default_rng(0)
e = rng.standard_normal(kit.glb_sz) + 1j * rng.standard_normal(kit.glb_sz)
s = mb.Scalar("FFF").upload(np.asfortranarray(e))
mb.chop(s)
if args.batch > 1:
    group = [s] + [s.copy() for _ in range(args.batch - 1)]
    for _ in range(args.reps):
        mb.trans_many(group, "PPP")
        mb.trans_many(group, "FFF")
else:
    for _ in range(args.reps):
        mb.trans(s, "PPP")
        mb.trans(s, "FFF")
mb.device_sync()
print("done", mb.launch_count())
