#!/usr/bin/env python
"""Host <-> device copy rates of one GPU from pinned memory: each direction alone and both at once (the ceiling of the
end-to-end figure of bench.py, whose timed region moves every field in and out over PCIe).

    python tools/pcie_duplex.py [--mb 256] [--reps 10]
"""
import argparse
import json

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=256)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
n = args.mb * (1 << 20)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d: bool, d2h: bool):
    for it in range(2 + args.reps):
        if it == 2:
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_event(e0)
            s2.wait_event(e0)
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) / args.reps * 1e-3) / 1e9


print(json.dumps({"h2d_GBps_alone": round(run(True, False), 1), "d2h_GBps_alone": round(run(False, True), 1),
                  "GBps_per_direction_both_at_once": round(run(True, True), 1), "mbytes": args.mb}))
