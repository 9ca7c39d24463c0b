"""GPU-box tool: probe the host, sweep pack-kernel variants/launch shapes on the
Llama-3-8B bf16 state_dict, and time the drain.  Writes gpurun_out/sweep.jsonl.
Not a bench (bench.py is); this is how launch defaults were chosen."""

import json
import os
import shutil
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlrover_b200 import _native as native  # noqa: E402
from dlrover_b200 import shapes  # noqa: E402

OUT = os.path.join("gpurun_out", "sweep.jsonl")
os.makedirs("gpurun_out", exist_ok=True)


def emit(**kw):
    with open(OUT, "a") as f:
        f.write(json.dumps(kw) + "\n")
    print(json.dumps(kw), flush=True)


def time_kernel(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    scale = float(os.environ.get("SWEEP_SCALE", "1.0"))
    du = shutil.disk_usage("/dev/shm")
    mem = {l.split(":")[0]: l.split(":")[1].strip() for l in open("/proc/meminfo") if ":" in l}
    emit(kind="host", cpus=os.cpu_count(), shm_total_gb=du.total / 1e9, shm_free_gb=du.free / 1e9,
         mem_total=mem.get("MemTotal"), mem_avail=mem.get("MemAvailable"),
         gpu=torch.cuda.get_device_name(0), gpus=torch.cuda.device_count())
    torch.cuda.set_device(0)
    ctx = native.get_context(0)
    sd = shapes.build_state_dict(shapes.scale_shapes(shapes.llama3_8b_shapes(), scale),
                                 torch.bfloat16, "cuda")
    leaves = list(sd.values())
    offs, off = [], 0
    for t in leaves:
        offs.append(off)
        off += t.numel() * t.element_size()
    S = off
    ctx.arena_reserve(S)
    emit(kind="workload", tensors=len(leaves), payload_bytes=S)
    stream = torch.cuda.current_stream()

    # torch baseline: one D2D copy of the same bytes
    a = torch.empty(S // 2, dtype=torch.bfloat16, device="cuda")
    b = torch.empty_like(a)
    med, best = time_kernel(lambda: b.copy_(a))
    emit(kind="torch_copy", ms=med, best_ms=best, gbs=2 * S / med / 1e6)
    del a, b

    for chunk in (64 << 10, 256 << 10, 1 << 20):
        plan = ctx.plan([t.data_ptr() for t in leaves], offs,
                        [t.numel() * t.element_size() for t in leaves], chunk)
        for cps in (2, 4, 6, 8):
            ctx.set_launch(lsu_ctas_per_sm=cps)
            med, best = time_kernel(lambda: plan.pack(stream, native.VARIANT_LSU))
            emit(kind="lsu", chunk=chunk, ctas_per_sm=cps, ms=med, best_ms=best,
                 gbs=2 * S / med / 1e6)
        for cps, stages, tile in ((1, 12, 16 << 10), (2, 6, 16 << 10), (2, 3, 32 << 10),
                                  (4, 3, 16 << 10), (4, 6, 8 << 10), (1, 6, 32 << 10),
                                  (3, 4, 16 << 10), (8, 3, 8 << 10)):
            ctx.set_launch(tma_ctas_per_sm=cps, tma_stages=stages, tma_tile_bytes=tile)
            med, best = time_kernel(lambda: plan.pack(stream, native.VARIANT_TMA))
            emit(kind="tma", chunk=chunk, ctas_per_sm=cps, stages=stages, tile=tile, ms=med,
                 best_ms=best, gbs=2 * S / med / 1e6)
        plan.destroy()

    # unpack + misaligned (shifted) path: offsets displaced by 4 bytes
    ctx.set_launch(lsu_ctas_per_sm=4)
    ctx.arena_reserve(S + 64)
    plan = ctx.plan([t.data_ptr() for t in leaves], [o + 4 for o in offs],
                    [t.numel() * t.element_size() for t in leaves], 256 << 10)
    med, best = time_kernel(lambda: plan.pack(stream, native.VARIANT_LSU))
    emit(kind="lsu_shift4", ms=med, best_ms=best, gbs=2 * S / med / 1e6)
    med, best = time_kernel(lambda: plan.unpack(stream, native.VARIANT_LSU))
    emit(kind="lsu_shift4_unpack", ms=med, best_ms=best, gbs=2 * S / med / 1e6)
    plan.destroy()

    # drain: pinned (cudaHostAlloc) vs registered /dev/shm mmap
    plan = ctx.plan([t.data_ptr() for t in leaves], offs,
                    [t.numel() * t.element_size() for t in leaves], 256 << 10)
    try:
        t0 = time.time()
        host = torch.empty(S, dtype=torch.uint8).pin_memory()
        emit(kind="pin_alloc", s=time.time() - t0)
        for i in range(3):
            tk = plan.save_async(host.data_ptr(), stream)
            ctx.save_wait(tk)
            p, d, tot = ctx.save_timings(tk)
            emit(kind="save_pinned", i=i, pack_ms=p, drain_ms=d, total_ms=tot, gbs=S / tot / 1e6)
        for i in range(2):
            for t in leaves[:4]:
                t.zero_()
            plan.restore_async(host.data_ptr(), stream)
            ctx.restore_wait()
            f, s, tot = ctx.restore_timings()
            emit(kind="restore_pinned", i=i, fill_ms=f, scatter_ms=s, total_ms=tot,
                 gbs=S / tot / 1e6)
        del host
    except Exception as e:  # noqa: BLE001
        emit(kind="save_pinned_error", err=str(e))

    import ctypes
    import mmap
    import _posixshmem
    name = f"/fc_sweep_{os.getpid()}"
    fd = _posixshmem.shm_open(name, os.O_CREAT | os.O_EXCL | os.O_RDWR, mode=0o600)
    try:
        os.ftruncate(fd, S)
        mm = mmap.mmap(fd, S)
        addr = ctypes.addressof(ctypes.c_char.from_buffer(mm))
        t0 = time.time()
        ctx.host_register(addr, S, prefault_threads=min(16, os.cpu_count() or 1))
        emit(kind="shm_register", s=time.time() - t0, threads=min(16, os.cpu_count() or 1))
        for i in range(3):
            tk = plan.save_async(addr, stream)
            ctx.save_wait(tk)
            p, d, tot = ctx.save_timings(tk)
            emit(kind="save_shm", i=i, pack_ms=p, drain_ms=d, total_ms=tot, gbs=S / tot / 1e6)
        ctx.host_unregister(addr)
    except Exception as e:  # noqa: BLE001
        emit(kind="save_shm_error", err=str(e))
    finally:
        _posixshmem.shm_unlink(name)
    plan.destroy()

    # the reference's way: per-tensor copy_ into pageable shm-like memory
    try:
        host = torch.empty(S, dtype=torch.uint8)
        host.zero_()
        for i in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for t, o in zip(leaves, offs):
                n = t.numel() * t.element_size()
                host[o:o + n].view(t.dtype).reshape(t.shape).copy_(t)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            emit(kind="ref_style_pageable_copy", i=i, s=dt, gbs=S / dt / 1e9)
    except Exception as e:  # noqa: BLE001
        emit(kind="ref_style_error", err=str(e))


if __name__ == "__main__":
    main()
