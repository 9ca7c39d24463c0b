"""Short view of a bench.py JSON line (argv[1]); on failure the tail of the stderr file (argv[2])."""
import json
import sys

try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    for k in ("n_gpus", "value", "e2e", "stall_ms", "restore", "restore_fresh_process",
              "ddp_cooperative", "fsdp", "drain"):
        v = d.get(k)
        if isinstance(v, dict):
            v = {a: b for a, b in v.items() if a not in ("method", "note", "api", "what")}
        print(k, json.dumps(v)[:1800])
except Exception as e:  # noqa: BLE001
    print("no bench line:", e)
    if len(sys.argv) > 2:
        print(open(sys.argv[2]).read()[-3500:])
