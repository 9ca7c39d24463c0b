"""N concurrent drains into tmpfs segments with a share of each segment's pages bound to the
OTHER socket (fc_host_bind_numa remote_per256).  Run under torchrun with the ranks' GPUs on
ONE socket (CUDA_VISIBLE_DEVICES=0,1,2,3 on the 8-GPU node: all four hang off NUMA node 0,
the other socket's memory controllers idle).  For each share: aggregate and per-rank GB/s of a
4 GiB drain per rank (best of 3, device-timed, all ranks start together)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DLROVER_LOG_LEVEL", "WARNING")
import torch
import torch.distributed as dist

rank, local, world = int(os.getenv("RANK", 0)), int(os.getenv("LOCAL_RANK", 0)), int(os.getenv("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from dlrover_b200 import _native as native
from dlrover_b200.common.multi_process import SharedMemory

SIZE = int(os.getenv("PROBE_GIB", "4")) << 30
ctx = native.get_context(local)
src = torch.empty(SIZE, dtype=torch.uint8, device=dev).random_(0, 255)
node, n_nodes = native.device_numa_node(local)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def gather(x):
    if world == 1:
        return [x]
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(v.item()) for v in out]


nodes = gather(float(node))
for share in [int(s) for s in os.getenv("SHARES", "0,43,64,85,107,128").split(",")]:
    shm = SharedMemory(name=f"numaprobe_{os.getppid()}_{rank}_{share}", create=True, size=SIZE)
    try:
        ctx.host_bind_numa(shm.address, SIZE, share)       # before the first touch
        os.environ["FC_NO_NUMA"] = "1"                      # host_register must not re-bind
        ctx.host_register(shm.address, SIZE, prefault_threads=16)
        os.environ.pop("FC_NO_NUMA", None)
        dst = torch.frombuffer(shm.buf, dtype=torch.uint8)
        stream, best = torch.cuda.Stream(), 0.0
        for _ in range(3):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record()
                for o in range(0, SIZE, 32 << 20):
                    dst[o:o + (32 << 20)].copy_(src[o:o + (32 << 20)], non_blocking=True)
                e1.record()
            e1.synchronize()
            best = max(best, SIZE / (e0.elapsed_time(e1) * 1e6))
        assert torch.equal(dst[-(1 << 20):].to(dev), src[-(1 << 20):])
        vals = gather(best)
        if rank == 0:
            print(json.dumps({"probe": "numa_split", "n_gpus": world, "gpu_numa_nodes": nodes,
                              "remote_per256": share, "remote_fraction": round(share / 256, 3),
                              "aggregate_GBps": round(sum(vals), 1),
                              "per_rank_GBps": [round(v, 1) for v in vals]}), flush=True)
        del dst
        ctx.host_unregister(shm.address)
    finally:
        shm.close()
        shm.unlink()
if world > 1:
    dist.destroy_process_group()
