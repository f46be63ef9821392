#!/bin/bash
# Runs ON THE GPU BOX: transform round-trip throughput sweep (BASELINE.json configs[4]) on one GPU, one JSON line per
# point in gpurun_out/sweep_<tag>.jsonl.  Field counts keep every point's batch larger than L2 and the run short.
TAG=${1:-r1}
OUT=gpurun_out/sweep_$TAG.jsonl
mkdir -p gpurun_out
: > $OUT
python bench.py --no-cpu --steps 10 --size 64 --fields 512 >> $OUT 2>> gpurun_out/sweep_$TAG.err
python bench.py --no-cpu --steps 10 --size 128 >> $OUT 2>> gpurun_out/sweep_$TAG.err
python bench.py --no-cpu --steps 5 --size 256 --fields 8 >> $OUT 2>> gpurun_out/sweep_$TAG.err
python bench.py --no-cpu --steps 5 --size 512 --fields 2 --batch 2 >> $OUT 2>> gpurun_out/sweep_$TAG.err
python bench.py --no-cpu --steps 3 --shape 1024,512,512 --fields 2 --batch 2 >> $OUT 2>> gpurun_out/sweep_$TAG.err
