"""Merge the rank*.json files of one tests/mgpu_worker.py run (argv[1]) into one JSON object."""
import glob
import json
import os
import sys

ranks = [json.load(open(f)) for f in sorted(glob.glob(os.path.join(sys.argv[1], "rank*.json")))]
for r in ranks:
    r.pop("traceback", None) if r.get("ok") else None
print(json.dumps({"ranks_reported": len(ranks), "all_ok": bool(ranks) and all(r.get("ok") for r in ranks),
                  "ranks": ranks}, default=str))
