"""BASELINE.json configs[0]: GPT-2 small (nanoGPT shapes, fp32, 652 MB incl. the tied
lm_head) on CPU, world_size 1: memory-save time of this repo's handler (fc_host_pack)
vs the REFERENCE's handler imported from /root/reference (only where it exists:
the build container) vs naive torch.save to /dev/shm."""
import json
import os
import sys
import time
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ROLE_NAME"] = "dlrover-trainer"
os.environ["TORCHELASTIC_RUN_ID"] = f"cfg0_{os.getpid()}"
os.environ["DLROVER_LOG_LEVEL"] = "ERROR"
import torch

from dlrover_b200 import shapes
from dlrover_b200.shm_handler import DLROVER_CKPT_CONFIG_KEY, CheckpointConfig, SharedMemoryHandler

sd = shapes.build_state_dict(shapes.gpt2_small_shapes(), torch.float32, "cpu")
S = shapes.payload_bytes(sd)
out = {"payload_bytes": S, "cores": os.cpu_count()}


def timeit(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2]


h = SharedMemoryHandler(0, host=True)
full = {"model_states": sd, DLROVER_CKPT_CONFIG_KEY: CheckpointConfig(step=1, paths={})}
t = timeit(lambda: h.save_state_dict(full))
out["ours_ms"] = t * 1e3
out["ours_GBps"] = S / t / 1e9
# restore into live CPU tensors: ours (fc_host_unpack, all cores) vs what a user of the
# reference does (load_state_dict -> views on the segment, then one copy_ per tensor)
target = {k: torch.empty_like(v) for k, v in sd.items()}
t = timeit(lambda: h.restore_into({"model_states": target}))
out["ours_restore_ms"] = t * 1e3
assert all(torch.equal(target[k], sd[k]) for k in list(sd)[:3] + list(sd)[-2:])


def ref_style_restore():
    views = h.load_state_dict()["model_states"]
    with torch.no_grad():
        for k, v in target.items():
            v.copy_(views[k])


t = timeit(ref_style_restore)
out["reference_style_restore_ms"] = t * 1e3
h.unlink()
h.close()

if os.path.isdir("/root/reference"):
    sys.path.insert(0, "/root/reference")
    sys.modules.setdefault("kubernetes", mock.MagicMock())
    logging_off = __import__("logging").disable(50)
    from dlrover.python.elastic_agent.torch import ckpt_saver as ref

    os.environ["TORCHELASTIC_RUN_ID"] = f"cfg0ref_{os.getpid()}"
    rh = ref.SharedMemoryHandler(0, host=True)
    rfull = {"model_states": sd, ref.DLROVER_CKPT_CONFIG_KEY: ref.CheckpointConfig(step=1, paths={})}
    t = timeit(lambda: rh.save_state_dict(rfull))
    out["reference_ms"] = t * 1e3
    out["reference_GBps"] = S / t / 1e9
    rh.shared_memory.unlink()

t = timeit(lambda: torch.save(sd, "/dev/shm/cfg0_naive.pt"), n=2)
os.remove("/dev/shm/cfg0_naive.pt")
out["torch_save_devshm_ms"] = t * 1e3
print(json.dumps(out))
