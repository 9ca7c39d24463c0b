"""cProfile of the training-thread side of save_checkpoint(MEMORY): where do the
~1.8 ms of host time per asynchronous save go?  Output: top functions by cumulative
time over N saves (each followed by wait_memory_save outside the profile)."""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TORCHELASTIC_RUN_ID", f"hprof{os.getppid()}")
os.environ.setdefault("DLROVER_LOG_LEVEL", "WARNING")
import torch

from dlrover_b200 import shapes
from dlrover_b200.flash_checkpoint.api import DdpCheckpointer, StorageType

N = int(os.getenv("N_SAVES", "40"))
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
ckpt = DdpCheckpointer(f"/tmp/fc_hprof_{os.getenv('TORCHELASTIC_RUN_ID')}")
sd = shapes.build_state_dict(shapes.llama3_8b_shapes(), torch.bfloat16, dev)
for s in range(1, 4):
    ckpt.save_checkpoint(s, sd, storage_type=StorageType.MEMORY)
    ckpt.wait_memory_save()
plain = []
for s in range(1000, 1000 + N):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ckpt.save_checkpoint(s, sd, storage_type=StorageType.MEMORY)
    plain.append(time.perf_counter() - t0)
    ckpt.wait_memory_save()
plain.sort()
print(f"saves={N} host call median {plain[N // 2] * 1e3:.3f} ms  min {plain[0] * 1e3:.3f} ms  "
      f"p90 {plain[int(N * 0.9)] * 1e3:.3f} ms (no profiler)")
prof = cProfile.Profile()
wall = []
for s in range(4, 4 + N):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prof.enable()
    ckpt.save_checkpoint(s, sd, storage_type=StorageType.MEMORY)
    prof.disable()
    wall.append(time.perf_counter() - t0)
    ckpt.wait_memory_save()
wall.sort()
print(f"saves={N} host call median {wall[N // 2] * 1e3:.3f} ms  min {wall[0] * 1e3:.3f} ms "
      f"(with cProfile overhead)")
out = io.StringIO()
pstats.Stats(prof, stream=out).sort_stats("cumulative").print_stats(45)
print(out.getvalue())
ckpt.engine.close()
