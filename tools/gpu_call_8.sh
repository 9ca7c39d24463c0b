#!/bin/bash
# launch list of OUR kernels during the bench command (the fill kernels of the synthetic state
# used up the launch cap of call 7)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fc_copy -c 300 --csv --log-file gpurun_out/c8_launches.csv python bench.py --steps 2 --warmup 3 --no-stall > gpurun_out/c8_bench_under_ncu.log 2>&1
grep -c "fc_copy" gpurun_out/c8_launches.csv
