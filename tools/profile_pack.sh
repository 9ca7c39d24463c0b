#!/bin/bash
# Run under gpurun (1 GPU).  Produces in gpurun_out/:
#   launches_fc.csv   every launch of our kernels during `bench.py` with its device time
#   prof_pack_tma.ncu-rep / prof_pack_lsu.ncu-rep   --set full captures of the gather kernel
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fc_copy \
    --csv --log-file gpurun_out/launches_fc.csv python bench.py --steps 2 --warmup 1 --no-stall \
    > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fc_copy_tma -s 3 -c 2 \
    -f -o gpurun_out/prof_pack_tma python bench.py --steps 2 --warmup 1 --no-stall \
    > gpurun_out/prof_tma.log 2>&1
DLROVER_B200_VARIANT=lsu timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:fc_copy_lsu -s 3 -c 2 -f -o gpurun_out/prof_pack_lsu python bench.py --steps 2 \
    --warmup 1 --no-stall > gpurun_out/prof_lsu.log 2>&1
ls -la gpurun_out
