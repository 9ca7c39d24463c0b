"""Agent-side persist: torch.save vs fast_torch_save.fast_save (same bytes) on a
host-resident state dict.  Needs no GPU; run on the GPU box because that is the
host whose cores/disks matter."""
import hashlib
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlrover_b200 import shapes
from dlrover_b200.fast_torch_save import fast_save


def sha(p):
    h = hashlib.sha256()
    with open(p, "rb") as f:
        for b in iter(lambda: f.read(1 << 24), b""):
            h.update(b)
    return h.hexdigest()


scale = float(os.getenv("PERSIST_SCALE", "0.25"))
sd = shapes.build_state_dict(shapes.scale_shapes(shapes.llama3_8b_shapes(), scale), torch.bfloat16,
                             "cpu")
S = shapes.payload_bytes(sd)
for root in ("/dev/shm", "/tmp"):
    d = os.path.join(root, f"fc_persist_{os.getpid()}")
    os.makedirs(d + "/a", exist_ok=True)
    os.makedirs(d + "/b", exist_ok=True)
    t0 = time.time()
    torch.save(sd, d + "/a/rank_0.pt")
    t1 = time.time()
    res = {"root": root, "bytes": S, "cores": os.cpu_count(),
           "torch_save_s": round(t1 - t0, 2), "torch_save_GBps": round(S / (t1 - t0) / 1e9, 2)}
    os.environ["DLROVER_B200_DIRECT_IO"] = "0"
    for th in (4, 16, 32):
        t0 = time.time()
        fast_save(sd, d + "/b/rank_0.pt", threads=th)
        dt = time.time() - t0
        res[f"fast_{th}t_s"] = round(dt, 2)
        res[f"fast_{th}t_GBps"] = round(S / dt / 1e9, 2)
    res["identical"] = sha(d + "/a/rank_0.pt") == sha(d + "/b/rank_0.pt")
    # O_DIRECT (second half): only meaningful on a block-device backed fs
    from dlrover_b200.common import direct_io
    os.environ["DLROVER_B200_DIRECT_IO"] = "1"
    res["fs_type"] = direct_io._fs_type(d)
    res["o_direct_usable"] = direct_io.enabled_for(d + "/b/rank_0.pt")
    if res["o_direct_usable"] and root != "/dev/shm":
        for th in (4, 16):
            t0 = time.time()
            fast_save(sd, d + "/b/rank_0.pt", threads=th)
            os.sync()
            dt = time.time() - t0
            res[f"direct_{th}t_incl_sync_s"] = round(dt, 2)
            res[f"direct_{th}t_GBps"] = round(S / dt / 1e9, 2)
        res["direct_identical"] = sha(d + "/a/rank_0.pt") == sha(d + "/b/rank_0.pt")
        os.environ["DLROVER_B200_DIRECT_IO"] = "0"
        t0 = time.time()
        fast_save(sd, d + "/b/rank_0.pt", threads=16)
        os.sync()
        dt = time.time() - t0
        res["buffered_16t_incl_sync_s"] = round(dt, 2)
        res["buffered_16t_incl_sync_GBps"] = round(S / dt / 1e9, 2)
    print(json.dumps(res), flush=True)
    os.remove(d + "/a/rank_0.pt")
    os.remove(d + "/b/rank_0.pt")
