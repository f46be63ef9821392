#!/bin/bash
# Launched by torchrun --no-python: rank 0 runs under ncu with a ONE-PASS metric set (a replayed kernel would repeat its
# side of the exchange barrier while the peers have moved on), the other ranks run plain.  NVLink byte counters of the
# exchange-bearing kernels on rank 0 (on the round-2 boxes the nvltx__/nvlrx__ counters fail with "UnknownError": the
# default NCU_METRICS=gpu__time_duration.sum gives the launch list, the bytes come from the exchange plan):
#   python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#       tools/ncu_rank0.sh OUT.csv bench.py --gpus 2 --steps 2 --warmup 3 --fields 16 --no-cpu --no-step
OUT=$1
shift
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --metrics ${NCU_METRICS:-gpu__time_duration.sum} \
    --clock-control none -k regex:'slab_ship|fft_reg_kernel|leg_backward_ws_kernel|exchange_put' \
    -c 120 --csv --log-file "$OUT" python "$@"
else
  exec python "$@"
fi
