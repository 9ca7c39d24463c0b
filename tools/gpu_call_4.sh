#!/bin/bash
# round 2, fourth GPU call (2 GPUs): the cases added since call 3 (fullshards, fsdp_full), bench N=2,
# host-call profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "fullshards or fsdp_full or coop or fsdp" > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
grep -E "passed|failed|FAILED|ERROR|rc=" gpurun_out/c4_pytest.log | tail -12
grep -A30 "AssertionError\|Error:" gpurun_out/c4_pytest.log | head -90
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c4_bench_n2.json 2> gpurun_out/c4_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c4_bench_n2.json").read().strip().split("\n")[-1])
    for k in ("value","e2e","stall_ms","restore","ddp_cooperative","fsdp"):
        v=d.get(k)
        if isinstance(v,dict): v={a:b for a,b in v.items() if a not in("method","note","api","what")}
        print(k, json.dumps(v)[:1500])
except Exception as e:
    print("bench n2 failed", e); print(open("gpurun_out/c4_bench_n2.err").read()[-3000:])
PY
CUDA_VISIBLE_DEVICES=0 N_SAVES=30 timeout 600 python tools/host_call_profile.py > gpurun_out/c4_host_call_profile.txt 2>&1
head -60 gpurun_out/c4_host_call_profile.txt
