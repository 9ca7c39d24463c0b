"""Restore of a 16 GB shard into live tensors: arena fill + scatter vs in-place H2D DMA."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TORCHELASTIC_RUN_ID", f"rprobe{os.getppid()}")
os.environ.setdefault("DLROVER_LOG_LEVEL", "WARNING")
import torch

from dlrover_b200 import shapes
from dlrover_b200.flash_checkpoint.api import DdpCheckpointer, StorageType

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
sd = shapes.build_state_dict(shapes.llama3_8b_shapes(), torch.bfloat16, dev)
S = shapes.payload_bytes(sd)
ckpt = DdpCheckpointer(f"/tmp/fc_rprobe_{os.getenv('TORCHELASTIC_RUN_ID')}")
ckpt.save_checkpoint(1, sd, storage_type=StorageType.MEMORY)
ckpt.wait_memory_save()
handler = ckpt.engine._shm_handler
for mode in ("arena", "direct", "arena", "direct"):
    os.environ["DLROVER_B200_RESTORE"] = mode
    for t in sd.values():
        t.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    stats = handler.restore_into({"model_states": sd})
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"mode": mode, "wall_ms": round(dt * 1e3, 1), "GBps": round(S / dt / 1e9, 1),
                      **{k: round(v, 1) for k, v in stats.items()}}), flush=True)
ckpt.engine.close()
