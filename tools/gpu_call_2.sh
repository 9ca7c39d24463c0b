#!/bin/bash
# round 2, second GPU call (1 GPU): whole GPU suite (no -x), pin probe with yielding slices,
# bench at three drain piece sizes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/c2_pytest.log | tail -15
rm -f gpurun_out/r02_probe.jsonl
timeout 600 python tools/r02_probe.py pin > gpurun_out/c2_probe_pin.log 2>&1
cat gpurun_out/r02_probe.jsonl
for mb in 32 24 16; do
  DLROVER_B200_DRAIN_PIECE_MB=$mb timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/c2_bench_piece$mb.json 2> gpurun_out/c2_bench_piece$mb.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c2_bench_piece$mb.json").read().strip().split("\n")[-1])
    print("piece $mb", "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "first_save_s", round(d["e2e"]["first_save_s"],2), "bg_pin_s", round(d["e2e"].get("background_pin_s",0),2), "stall", {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["stall_ms"].items() if k in ("async","blocking","host_call_ms_async","pack_kernel_ms")}, "restore", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["restore"].items() if k!="device_times"})
except Exception as e:
    print("piece $mb failed", e); print(open("gpurun_out/c2_bench_piece$mb.err").read()[-1500:])
PY
done
