#!/bin/bash
# round 2, second 8-GPU call: BASELINE configs[3]/[4] shapes again, timed in the steady state
# (after the background pin), + the cooperative DDP leg with equal windows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
rm -rf /tmp/mg8_zero3; mkdir -p /tmp/mg8_zero3
timeout 500 bash -c "$(declare -f run); run 8 29625 tests/mgpu_worker.py --case zero3 --out /tmp/mg8_zero3 --flat-mib 25200 --in-place --snapshot-mib 32768 --full-compare 0" > $O/c6_zero3.log 2>&1
echo "zero3 rc=$?"; python tools/print_worker.py /tmp/mg8_zero3 > $O/c6_zero3.json; head -c 1800 $O/c6_zero3.json; echo
rm -f /dev/shm/mg* 2>/dev/null
rm -rf /tmp/mg8_megatron; mkdir -p /tmp/mg8_megatron
timeout 400 bash -c "$(declare -f run); run 8 29626 tests/mgpu_worker.py --case megatron --out /tmp/mg8_megatron --widths 1.0 --layers 8 --full-compare 0" > $O/c6_megatron.log 2>&1
echo "megatron rc=$?"; python tools/print_worker.py /tmp/mg8_megatron > $O/c6_megatron.json; head -c 1800 $O/c6_megatron.json; echo
rm -f /dev/shm/mg* 2>/dev/null
BENCH_ONLY=coop timeout 300 bash -c "$(declare -f run); run 8 29627 bench.py --gpus 8 --steps 4 --warmup 3 --no-stall" > $O/c6_bench_coop_n8.json 2> $O/c6_bench_coop_n8.err
python tools/print_bench.py $O/c6_bench_coop_n8.json $O/c6_bench_coop_n8.err
