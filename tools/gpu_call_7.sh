#!/bin/bash
# round 2, final single-GPU call: whole GPU suite, smoke, bench at two drain piece sizes,
# persist bench (O_DIRECT), ncu launch list of the bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/c7_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c7_pytest.log
grep -E "passed|failed|FAILED|ERROR|rc=" $O/c7_pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c7_smoke.log 2>&1; tail -2 $O/c7_smoke.log
for mb in 16 32; do
  DLROVER_B200_DRAIN_PIECE_MB=$mb timeout 700 python bench.py --steps 8 --warmup 3 > $O/c7_bench_piece$mb.json 2> $O/c7_bench_piece$mb.err
  echo "== piece $mb MiB"; python tools/print_bench.py $O/c7_bench_piece$mb.json $O/c7_bench_piece$mb.err | cut -c1-900
done
PERSIST_SCALE=0.25 timeout 600 python tools/persist_bench.py > $O/c7_persist.jsonl 2> $O/c7_persist.err; cat $O/c7_persist.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c7_launches.csv python bench.py --steps 2 --warmup 3 --no-stall > $O/c7_bench_under_ncu.log 2>&1
grep -c "fc_copy" $O/c7_launches.csv
