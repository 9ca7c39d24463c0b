"""ncu target: hybrid save of an AdamW-shaped state — a 4-byte `step` scalar first (drained
in place), then the Llama-3-8B tensors (16.06 GB) snapshotted from cut = 4: every range is
NOT congruent mod 16 with its source, the case the LSU slice served at 0.81 of peak with
1.16x read traffic in round 1.  Round 2 routes the slice through the bulk/shift/resid tables."""
import ctypes
import mmap
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _posixshmem  # noqa: E402
from dlrover_b200 import _native as native  # noqa: E402
from dlrover_b200 import shapes  # noqa: E402

torch.cuda.set_device(0)
ctx = native.get_context(0)
sd = shapes.build_state_dict(shapes.llama3_8b_shapes(), torch.bfloat16, "cuda", fill=False)
leaves = list(sd.values())
step = torch.ones(1, device="cuda")
offs, o = [0], 4
for t in leaves:
    offs.append(o)
    o += t.numel() * 2
plan = ctx.plan([step.data_ptr()] + [t.data_ptr() for t in leaves], offs,
                [4] + [t.numel() * 2 for t in leaves])
ctx.arena_reserve(o)
name = f"/fc_ncu_hybrid_{os.getpid()}"
fd = _posixshmem.shm_open(name, os.O_CREAT | os.O_EXCL | os.O_RDWR, mode=0o600)
os.ftruncate(fd, o)
mm = mmap.mmap(fd, o)
addr = ctypes.addressof(ctypes.c_char.from_buffer(mm))
ctx.host_register(addr, o, prefault_threads=16)
s = torch.cuda.current_stream()
for _ in range(int(os.getenv("REPS", "3"))):
    tk = plan.save_hybrid_async(addr, 4, s)
    ctx.save_wait(tk)
    print("gather_ms", ctx.save_timings(tk)[0])
plan.destroy()
ctx.host_unregister(addr)
mm.close()
os.close(fd)
_posixshmem.shm_unlink(name)
