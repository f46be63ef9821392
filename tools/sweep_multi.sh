#!/bin/bash
# Runs ON THE GPU BOX (gpurun --gpus N): transform round-trip sweep (BASELINE.json configs[4]) with every field
# slab-distributed over N GPUs -- the field shape of tools/sweep.sh's one-GPU points, N times its batch (--weak fields:
# the work per GPU stays that of the one-GPU point).  One JSON line per point in gpurun_out/sweep_<tag>_<N>gpu.jsonl.
#   tools/sweep_multi.sh N TAG [points...]     points: 128 256 512 1024 (default: all)
N=${1:-2}
TAG=${2:-r2}
shift 2
POINTS=${@:-128 256 512 1024}
OUT=gpurun_out/sweep_${TAG}_${N}gpu.jsonl
ERR=gpurun_out/sweep_${TAG}_${N}gpu.err
mkdir -p gpurun_out
: > $OUT
run() {
  local t=$1; shift
  timeout $t python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N --no-cpu --no-step --weak fields "$@" 2>> $ERR | grep '^{' >> $OUT
}
for p in $POINTS; do
  case $p in
    128) run 300 --steps 10 --size 128 ;;
    256) run 300 --steps 5 --size 256 --fields 8 ;;
    512) run 400 --steps 5 --size 512 --fields 2 --batch 2 ;;
    1024) run 600 --steps 3 --shape 1024,512,512 --fields 2 --batch 2 ;;
  esac
done
python -c "
import json
for l in open('$OUT'):
    d = json.loads(l)
    print(d['config']['workload'][:48], 'N=%d' % d['n_gpus'], '%.1f GDOF/s' % d['value'], '%.3f ms/step' % d['ms_per_step'])
"
