#!/bin/bash
# Runs ON THE GPU BOX (under gpurun, ONE GPU): ncu launch list of the bench command, full captures of one round trip
# (the six transform kernels) and of one q-vortex time step (every kernel class of the step).
# Reports are written to /tmp (gpurun returns at most 64 MiB); the summaries (tools/ncu_summary.py) and the reports
# small enough to travel go to gpurun_out/ and from there, by hand, to profiles/.
#   tools/collect_profiles.sh TAG [parts...]     parts: launches trans128 trans512 step256 (default: all)
set -x
mkdir -p gpurun_out /tmp/ncu
TAG=${1:-r2}
shift
PARTS=${@:-launches trans128 trans512 step256}
for part in $PARTS; do
  case $part in
    launches)
      # every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches128_$TAG.csv \
        python bench.py --steps 2 --warmup 3 --fields 16 --no-cpu --no-step > gpurun_out/launches128_$TAG.log 2>&1 ;;
    trans128)
      # one batched round trip at 128^3 (8 scalars per launch, the bench's launch shape): all six kernels, full sets
      # (skip the first round trip)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fft_reg|leg_' -s 6 -c 6 -o /tmp/ncu/trans128_full_$TAG \
        python tools/prof_roundtrip.py --size 128 --reps 3 --batch 8 > gpurun_out/ncu128_$TAG.log 2>&1 ;;
    trans512)
      # the north-star size, one scalar per launch
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fft_reg|leg_' -s 6 -c 6 -o /tmp/ncu/trans512_full_$TAG \
        python tools/prof_roundtrip.py --size 512 --reps 2 > gpurun_out/ncu512_$TAG.log 2>&1 ;;
    step256)
      # one q-vortex ABCN step at 256^3 (input.params physics): every launch of the step, full sets
      timeout 900 ncu --set full --clock-control none --profile-from-start off -c 90 -o /tmp/ncu/step256_full_$TAG \
        python tools/step_bench.py --size 256 --steps 2 --warmup 1 --ncu-step > gpurun_out/ncustep256_$TAG.log 2>&1 ;;
  esac
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_after.csv
for f in trans128_full_$TAG trans512_full_$TAG step256_full_$TAG; do
  [ -f /tmp/ncu/$f.ncu-rep ] || continue
  python tools/ncu_summary.py /tmp/ncu/$f.ncu-rep gpurun_out/${f}_summary.csv
  sz=$(stat -c %s /tmp/ncu/$f.ncu-rep)
  if [ "$sz" -lt 20000000 ]; then cp /tmp/ncu/$f.ncu-rep gpurun_out/; fi
done
du -sh gpurun_out
