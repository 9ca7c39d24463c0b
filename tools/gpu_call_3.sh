#!/bin/bash
# round 2, third GPU call (2 GPUs): multi-rank parity under NCCL, bench N=2 (both arms),
# plus the single-GPU bench with the sliced-pin fix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c3_topo.txt 2>&1
timeout 1800 python -m pytest tests/test_multi_gpu.py tests/test_gpu_r02.py -m gpu -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
grep -E "passed|failed|FAILED|ERROR|rc=" gpurun_out/c3_pytest.log | tail -12
grep -B2 -A25 "Error\|assert" gpurun_out/c3_pytest.log | head -80
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c3_bench_n2.json 2> gpurun_out/c3_bench_n2.err
tail -c 3000 gpurun_out/c3_bench_n2.json; tail -5 gpurun_out/c3_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/c3_bench_ref_n2.json 2> gpurun_out/c3_bench_ref_n2.err
tail -c 600 gpurun_out/c3_bench_ref_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c3_bench_n1.json 2> gpurun_out/c3_bench_n1.err
tail -c 2500 gpurun_out/c3_bench_n1.json; tail -3 gpurun_out/c3_bench_n1.err
