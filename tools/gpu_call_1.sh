#!/bin/bash
# round 2, first GPU call (1 GPU): tests, probes, ncu of the hybrid snapshot, a short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_env.txt; free -g >> gpurun_out/c1_env.txt; df -h /dev/shm >> gpurun_out/c1_env.txt; nproc >> gpurun_out/c1_env.txt
lsblk >> gpurun_out/c1_env.txt 2>&1; mount | grep -E " / | /tmp " >> gpurun_out/c1_env.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
rm -f gpurun_out/r02_probe.jsonl
timeout 600 python tools/r02_probe.py hybrid > gpurun_out/c1_probe_hybrid.log 2>&1
timeout 900 python tools/r02_probe.py drain > gpurun_out/c1_probe_drain.log 2>&1
timeout 900 python tools/r02_probe.py staged > gpurun_out/c1_probe_staged.log 2>&1
timeout 600 python tools/r02_probe.py pin > gpurun_out/c1_probe_pin.log 2>&1
gcc -O2 -mavx2 -pthread -o /tmp/mem_write_bw tools/csrc/mem_write_bw.c && timeout 300 /tmp/mem_write_bw 16 512 3 > gpurun_out/c1_mem_write_bw.jsonl 2>&1
timeout 300 /tmp/mem_write_bw 32 256 3 >> gpurun_out/c1_mem_write_bw.jsonl 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_copy -c 6 -o gpurun_out/c1_hybrid_ncu python tools/ncu_hybrid.py > gpurun_out/c1_ncu_hybrid.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_target.py > gpurun_out/c1_sanitizer_memcheck.log 2>&1
tail -3 gpurun_out/c1_sanitizer_memcheck.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench_n1.json 2> gpurun_out/c1_bench_n1.err
tail -c 600 gpurun_out/c1_bench_n1.json
cat gpurun_out/r02_probe.jsonl | tail -40
