#!/bin/bash
# round 2, 8-GPU call: bench N=8, NUMA split probe + bench N=4 on one socket's GPUs, 8-rank parity
# cases, BASELINE configs[3]/[4] at the scale host RAM allows.  Every step has its own timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi -L; nvidia-smi topo -m; free -g; df -h /dev/shm; nproc; } > $O/c5_env.txt 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }

# 1. bench, 8 ranks (weak-scaling value + cooperative DDP + FSDP legs)
timeout 900 bash -c "$(declare -f run); run 8 29621 bench.py --gpus 8 --steps 4 --warmup 3 --no-stall" > $O/c5_bench_n8.json 2> $O/c5_bench_n8.err
python tools/print_bench.py $O/c5_bench_n8.json $O/c5_bench_n8.err
ls /dev/shm | head -5; rm -f /dev/shm/fcbench* 2>/dev/null

# 2. GPUs 0-3 (one socket): where should the pages go?
CUDA_VISIBLE_DEVICES=0,1,2,3 SHARES=0,64,85,107 timeout 400 bash -c "$(declare -f run); run 4 29622 tools/numa_split_probe.py" > $O/c5_numa_split_n4.jsonl 2> $O/c5_numa_split_n4.err
cat $O/c5_numa_split_n4.jsonl
BEST=$(python - <<'PY'
import json
best, share = 0, 0
for line in open("gpurun_out/c5_numa_split_n4.jsonl"):
    try:
        d = json.loads(line)
    except Exception:
        continue
    if d.get("probe") == "numa_split" and d["aggregate_GBps"] > best * 1.02:
        best, share = d["aggregate_GBps"], d["remote_per256"]
print(share)
PY
)
echo "best remote share per 256: $BEST"
CUDA_VISIBLE_DEVICES=0,1,2,3 DLROVER_B200_NUMA_REMOTE_PER256=$BEST BENCH_NO_FSDP=1 timeout 700 bash -c "$(declare -f run); run 4 29623 bench.py --gpus 4 --steps 4 --warmup 3 --no-stall" > $O/c5_bench_n4_share$BEST.json 2> $O/c5_bench_n4.err
python tools/print_bench.py $O/c5_bench_n4_share$BEST.json $O/c5_bench_n4.err
rm -f /dev/shm/fcbench* 2>/dev/null

# 3. parity at 8 ranks (every rank: whole segment == oracle image, restore)
for c in coop fsdp fullshards fsdp_full; do
  rm -rf /tmp/mg8_$c; mkdir -p /tmp/mg8_$c
  timeout 400 bash -c "$(declare -f run); run 8 29624 tests/mgpu_worker.py --case $c --out /tmp/mg8_$c --scale 0.03125" > $O/c5_mg8_$c.log 2>&1
  echo "case $c rc=$?"
  python tools/print_worker.py /tmp/mg8_$c > $O/c5_mg8_$c.json; cat $O/c5_mg8_$c.json | head -c 1500; echo
done

# 4. configs[3]: ZeRO-3 flat fp32 partitions, 3 x 26.4 GB per rank (0.75 of the 70B/8 shape), hybrid in-place
rm -rf /tmp/mg8_zero3; mkdir -p /tmp/mg8_zero3
timeout 700 bash -c "$(declare -f run); run 8 29625 tests/mgpu_worker.py --case zero3 --out /tmp/mg8_zero3 --flat-mib 25200 --in-place --snapshot-mib 32768 --full-compare 0" > $O/c5_zero3.log 2>&1
echo "zero3 rc=$?"; python tools/print_worker.py /tmp/mg8_zero3 > $O/c5_zero3.json; head -c 2500 $O/c5_zero3.json; echo
rm -f /dev/shm/mg* 2>/dev/null

# 5. configs[4]: Mixtral TP2xPP2 shards at full widths, 8 of 16 layers per stage (0.5 of the shape)
rm -rf /tmp/mg8_megatron; mkdir -p /tmp/mg8_megatron
timeout 600 bash -c "$(declare -f run); run 8 29626 tests/mgpu_worker.py --case megatron --out /tmp/mg8_megatron --widths 1.0 --layers 8 --full-compare 0" > $O/c5_megatron.log 2>&1
echo "megatron rc=$?"; python tools/print_worker.py /tmp/mg8_megatron > $O/c5_megatron.json; head -c 2500 $O/c5_megatron.json; echo
rm -f /dev/shm/mg* 2>/dev/null
free -g | head -2
