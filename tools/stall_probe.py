"""Per-step timing of a synthetic training loop around async checkpoints, to see
WHERE the stall lands.  Run under torchrun (N>=1).  Variants via env:
  PROBE_SYNC=item|stream     how the step syncs (D2H .item() vs stream sync)
  PROBE_NOREADY=1            skip the NCCL readiness all-reduce
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TORCHELASTIC_RUN_ID", f"probe{os.getppid()}")
os.environ.setdefault("DLROVER_LOG_LEVEL", "WARNING")
import torch
import torch.distributed as dist

rank, local, world = int(os.getenv("RANK", 0)), int(os.getenv("LOCAL_RANK", 0)), int(os.getenv("WORLD_SIZE", 1))
if world > 1:
    os.environ.setdefault("LOCAL_WORLD_SIZE", str(world))
    dist.init_process_group("nccl")
from dlrover_b200 import shapes
from dlrover_b200.flash_checkpoint import engine as eng_mod
from dlrover_b200.flash_checkpoint.api import DdpCheckpointer, StorageType

if os.getenv("PROBE_NOREADY") == "1":
    eng_mod.check_all_rank_ready = lambda group, ready: ready
ckpt = DdpCheckpointer(f"/tmp/fc_probe_{os.getenv('TORCHELASTIC_RUN_ID')}", local_shard_num=world,
                       global_shard_num=world)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
sd = shapes.build_state_dict(shapes.llama3_8b_shapes(), torch.bfloat16, dev)
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16) * 0.01
mode = os.getenv("PROBE_SYNC", "item")


def step():
    c = a
    for _ in range(40):
        c = torch.mm(c, b)
    if mode == "item":
        return float(c[0, 0].item())
    torch.cuda.current_stream().synchronize()


overlap = os.getenv("PROBE_OVERLAP") == "1"
if overlap:
    ckpt.engine.snapshot_stream = torch.cuda.Stream()
ckpt.save_checkpoint(1, sd, storage_type=StorageType.MEMORY)
ckpt.wait_memory_save()
for _ in range(5):
    step()
times, marks = [], []
for i in range(60):
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    if overlap and i % 20 == 11:
        # "optimizer.step()" of the step after the checkpoint: first mutation
        torch.cuda.current_stream().wait_event(ckpt.engine.pack_done_event())
    if i % 20 == 10:
        ckpt.save_checkpoint(100 + i, sd, storage_type=StorageType.MEMORY)
        marks.append(i)
    t2 = time.perf_counter()
    times.append((round((t1 - t0) * 1e3, 1), round((t2 - t1) * 1e3, 1)))
ckpt.wait_memory_save()
if rank == 0:
    print(json.dumps({"world": world, "sync": mode, "overlap": overlap, "noready": os.getenv("PROBE_NOREADY"),
                      "save_at": marks, "step_ms,save_call_ms": times}), flush=True)
if world > 1:
    dist.barrier()
ckpt.engine.close()
