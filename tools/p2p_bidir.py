#!/usr/bin/env python
"""Peer-copy rate between GPU 0 and GPU 1 of one box, one direction and both directions at once (copy engines,
cudaMemcpyPeerAsync through torch): the ceiling the exchange kernels' NVLink figures are compared with.  An all-to-all
sends and receives at the same time, so the both-directions figure is the relevant one.

    python tools/p2p_bidir.py [--mb 256] [--reps 20]
"""
import argparse
import json

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=256)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
n = args.mb * (1 << 20)
a0 = torch.empty(n, dtype=torch.uint8, device="cuda:0")
b0 = torch.empty(n, dtype=torch.uint8, device="cuda:0")
a1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
b1 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
s0 = torch.cuda.Stream(device="cuda:0")
s1 = torch.cuda.Stream(device="cuda:1")


def run(both: bool):
    for it in range(3 + args.reps):
        if it == 3:
            torch.cuda.synchronize("cuda:0")
            torch.cuda.synchronize("cuda:1")
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s0):
                e0.record()
        with torch.cuda.stream(s0):
            b1.copy_(a0, non_blocking=True)          # 0 -> 1, issued by GPU 0
        if both:
            with torch.cuda.stream(s1):
                b0.copy_(a1, non_blocking=True)      # 1 -> 0, issued by GPU 1
    torch.cuda.synchronize("cuda:1")
    with torch.cuda.stream(s0):
        e1.record()
    torch.cuda.synchronize("cuda:0")
    ms = e0.elapsed_time(e1) / args.reps
    return n / (ms * 1e-3) / 1e9


uni = run(False)
bi = run(True)
print(json.dumps({"peer_copy_GBps_one_direction": round(uni, 1), "peer_copy_GBps_per_direction_both_at_once": round(bi, 1),
                  "mbytes": args.mb, "reps": args.reps}))
