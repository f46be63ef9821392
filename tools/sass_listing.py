#!/usr/bin/env python
"""Writes profiles/<round>/sass_legendre.txt: mnemonic counts of the in-tree library and excerpts of the production
Legendre kernels (cuobjdump -sass; no GPU needed).

    python tools/sass_listing.py profiles/r2/sass_legendre.txt
"""
import re
import subprocess
import sys

LIB = "mlegs_b200/lib/libmlegs_b200.so"
out = sys.argv[1] if len(sys.argv) > 1 else "profiles/r2/sass_legendre.txt"
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())
allins = [i for v in funcs.values() for i in v]


def count(pat, ins):
    return sum(1 for i in ins if re.search(pat, i))


with open(out, "w") as fh:
    fh.write("# cuobjdump -sass of %s (sm_100a): the production Legendre kernels (tools/sass_listing.py)\n" % LIB)
    fh.write("# mnemonic counts over the whole library:\n")
    for pat in ["DMMA", "UBLKCP", "UTMALDG", "UTMASTG|UBLKRED|UBLKCP.G.S", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK",
                "USETMAXREG", "STG.E.128|ST.E.128", "LDG.E.128|LD.E.128", "UTCHMMA|UTCQMMA|UTCIMMA"]:
        fh.write("#   %s: %d\n" % (pat, count(pat, allins)))
    fh.write("# (no UTC*MMA: tcgen05 has no f64 kind; the FP64 tensor instruction is DMMA.8x8x4)\n")
    for name, ins in funcs.items():
        if "leg_forward_ws_kernel" not in name and "leg_backward_ws_kernel" not in name:
            continue
        fh.write("\n==== %s\n" % name)
        fh.write("# instructions: %d  DMMA: %d  LDS: %d  UBLKCP: %d  UTMALDG: %d  SYNCS: %d  BAR: %d  USETMAXREG: %d\n" % (
            len(ins), count("DMMA", ins), count(r"\bLDS", ins), count("UBLKCP", ins), count("UTMALDG", ins),
            count("SYNCS", ins), count(r"\bBAR", ins), count("USETMAXREG", ins)))
        fh.write("# TMA issue, mbarrier traffic and register re-allocation:\n")
        for i in ins:
            if re.search(r"UBLKCP|UTMALDG|SYNCS.ARRIVE.TRANS64|USETMAXREG", i):
                fh.write(i[:110] + "\n")
        first = next((k for k, i in enumerate(ins) if "DMMA" in i), None)
        if first is not None:
            fh.write("# the instruction stream around the first DMMAs of a contraction body:\n")
            for i in ins[max(0, first - 12):first + 36]:
                fh.write(i[:110] + "\n")
print("wrote", out)
