"""world_size=2 over gloo on CPU: the N>1 host logic — saving-rank policy,
readiness collective, per-rank segments, barrier + SAVE event, done-file commit
with global_shard_num=2, step-consistency check on load
(reference: checkpoint_egine_test.py:63-76,331-342 uses mp.spawn the same way)."""

import os
import sys
import tempfile
import time
import uuid

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, run_id, ckpt_dir, mode, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update({
        "RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world),
        "LOCAL_WORLD_SIZE": str(world), "GROUP_WORLD_SIZE": "1", "GROUP_RANK": "0",
        "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "TORCHELASTIC_RUN_ID": run_id,
        "DLROVER_LOG_LEVEL": "WARNING",
    })
    os.environ.pop("ROLE_NAME", None)  # local rank 0 forks the saver daemon
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dlrover_b200.common.constants import CheckpointConstant
    from dlrover_b200.common.storage import PosixDiskStorage
    from dlrover_b200.flash_checkpoint.api import DdpCheckpointer, StorageType
    from dlrover_b200.flash_checkpoint.engine import check_all_rank_ready

    result = {}
    try:
        if mode == "sharded":
            ckpt = DdpCheckpointer(ckpt_dir, local_shard_num=world, global_shard_num=world)
            eng = ckpt.engine
            result["saving_ranks"] = eng._saving_ranks
            result["shard_id"] = eng.local_shard_id
            sd = {"w": torch.full((1000,), float(rank + 1)), "rank": rank}
            ckpt.save_checkpoint(3, sd, storage_type=StorageType.MEMORY)
            back = ckpt.load_checkpoint()
            result["mem_ok"] = bool(torch.equal(back["w"], sd["w"]) and back["rank"] == rank)
            del back
            ckpt.save_checkpoint(4, sd, storage_type=StorageType.DISK)
            ckpt.wait_latest_checkpoint(timeout=90)
            result["tracker"] = open(os.path.join(ckpt_dir, "dlrover_latest.txt")).read()
            result["files"] = sorted(os.listdir(os.path.join(ckpt_dir, "4")))
            mine = torch.load(os.path.join(ckpt_dir, "4", f"rank_{rank}.pt"))
            result["disk_ok"] = bool(torch.equal(mine["w"], sd["w"]))
            result["ready_all"] = bool(check_all_rank_ready(None, True))
            result["ready_one_not"] = bool(check_all_rank_ready(None, rank != 1))
        elif mode == "rank0_only":
            # legal with the reference: only rank 0 calls save_checkpoint(MEMORY), the other
            # local ranks never show up -> the leader saves alone (after one short wait)
            os.environ["DLROVER_B200_COOP_JOIN_TIMEOUT_S"] = "0.5"
            ckpt = DdpCheckpointer(ckpt_dir)
            eng = ckpt.engine
            sd = {"w": torch.arange(3_000_000, dtype=torch.float32)}
            if rank == 0:
                t0 = time.time()
                ckpt.save_checkpoint(1, sd, storage_type=StorageType.MEMORY)
                ckpt.wait_memory_save()
                result["first_s"] = time.time() - t0
                sd["w"].add_(1)
                t0 = time.time()
                ckpt.save_checkpoint(2, sd, storage_type=StorageType.MEMORY)
                ckpt.wait_memory_save()
                result["second_s"] = time.time() - t0
                result["coop_wanted_after"] = eng._coop_wanted
            dist.barrier()
            back = ckpt.load_checkpoint()      # every rank reads the node's one image
            result["load_ok"] = bool(torch.equal(back["w"], torch.arange(
                3_000_000, dtype=torch.float32) + 1))
            del back
        elif mode == "cooperative":
            # replicated state, every local rank writes its slice of the ONE image
            ckpt = DdpCheckpointer(ckpt_dir)
            eng = ckpt.engine
            result["coop"] = bool(eng._cooperative())
            g = torch.Generator().manual_seed(5)
            sd = {"a": torch.randn(700_001, generator=g), "step": 3,
                  "b": [torch.arange(1_500_000, dtype=torch.int32), torch.ones(3)],
                  "c": torch.randn(2_000_003, generator=g).to(torch.bfloat16)}
            want = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in sd.items() if k != "b"}
            for step in (5, 6):      # second save: steady state (meta tree unchanged)
                sd["a"].add_(1.0)
                want["a"] = sd["a"].clone()
                ckpt.save_checkpoint(step, sd, storage_type=StorageType.MEMORY)
                ckpt.wait_memory_save()
                dist.barrier()
                back = ckpt.load_checkpoint()
                result[f"mem_ok_{step}"] = bool(
                    torch.equal(back["a"], want["a"]) and torch.equal(back["c"], want["c"])
                    and torch.equal(back["b"][0], sd["b"][0]) and back["step"] == 3)
                del back
                dist.barrier()
            plane = eng._shm_handler.metadata
            result["dict_sets"], result["ctl_publishes"] = plane.dict_sets, plane.ctl_publishes
            shm = eng._shm_handler.shared_memory
            result["segment"] = shm.name
            total = shm.size
            ctx = __import__("dlrover_b200.shm_handler", fromlist=["CoopContext"]).CoopContext(
                None, rank, world, 0)
            result["window"] = ctx.window(total)
            result["total"] = total
            ckpt.save_checkpoint(7, sd, storage_type=StorageType.DISK)
            if rank == 0:
                ckpt.wait_latest_checkpoint(timeout=90)
            dist.barrier()
            result["files"] = sorted(os.listdir(os.path.join(ckpt_dir, "7")))
            mine = torch.load(os.path.join(ckpt_dir, "7", "rank_0.pt"))
            result["disk_ok"] = bool(torch.equal(mine["a"], sd["a"]) and
                                     torch.equal(mine["c"], sd["c"]))
            # in-place restore into live tensors on every rank
            live = {k: (torch.zeros_like(v) if torch.is_tensor(v) else v) for k, v in sd.items()
                    if k != "b"}
            live["b"] = [torch.zeros_like(sd["b"][0]), torch.zeros(3)]
            result["restored_step"] = ckpt.load_checkpoint_into(live)
            result["restore_ok"] = bool(torch.equal(live["a"], sd["a"]) and
                                        torch.equal(live["b"][0], sd["b"][0]))
        else:  # replicated, the reference's policy: only local rank 0 writes
            os.environ["DLROVER_B200_COOP_DRAIN"] = "0"
            ckpt = DdpCheckpointer(ckpt_dir)
            eng = ckpt.engine
            result["saving_ranks"] = eng._saving_ranks
            sd = {"w": torch.arange(100, dtype=torch.float32)}
            ckpt.save_checkpoint(7, sd, storage_type=StorageType.DISK)
            result["cached_step"] = eng._cached_step
            # only a saving rank tracks latest_step; the others would poll
            # until the timeout (same in the reference)
            if rank == 0:
                ckpt.wait_latest_checkpoint(timeout=90)
            dist.barrier()
            result["files"] = sorted(os.listdir(os.path.join(ckpt_dir, "7")))
            # rank 1 never wrote memory: steps differ across ranks, so nobody
            # restores from memory and both fall back to rank 0's file
            loaded = ckpt.load_checkpoint()
            result["load_ok"] = bool(torch.equal(loaded["w"], sd["w"]))
        dist.barrier()
        eng.close()
    finally:
        torch.save(result, os.path.join(out_dir, f"r{rank}.pt"))
        dist.destroy_process_group()


def _run(mode):
    world = 2
    run_id = "mp" + uuid.uuid4().hex[:8]
    port = 29600 + (os.getpid() * 7 + hash(mode)) % 1500
    with tempfile.TemporaryDirectory() as ckpt_dir, tempfile.TemporaryDirectory() as out_dir:
        mp.spawn(_worker, args=(world, port, run_id, ckpt_dir, mode, out_dir), nprocs=world,
                 join=True)
        results = [torch.load(os.path.join(out_dir, f"r{r}.pt")) for r in range(world)]
    import glob
    import shutil
    shutil.rmtree(os.path.join("/tmp/ckpt_sock", run_id), ignore_errors=True)
    for f in glob.glob(f"/dev/shm/{run_id}_*"):
        os.unlink(f)
    return results


@pytest.mark.timeout(300)
def test_every_rank_is_a_shard():
    r0, r1 = _run("sharded")
    for r, res in enumerate((r0, r1)):
        assert res["saving_ranks"] == [0, 1]
        assert res["shard_id"] == r
        assert res["mem_ok"] and res["disk_ok"]
        assert res["tracker"] == "4"
        assert res["files"] == ["rank_0.pt", "rank_1.pt"]
        assert res["ready_all"] is True and res["ready_one_not"] is False


@pytest.mark.timeout(300)
def test_replicated_state_is_saved_cooperatively():
    """Default for replicated state on a multi-rank node: both ranks write their slice of
    the one image into ckpt_shm_0; one meta, one file on disk, identical to what a single
    saving rank would have produced (the tensors come back bit-exact on both ranks)."""
    r0, r1 = _run("cooperative")
    for res in (r0, r1):
        assert res["coop"] is True
        assert res["mem_ok_5"] and res["mem_ok_6"] and res["disk_ok"] and res["restore_ok"]
        assert res["restored_step"] == 7
        assert res["files"] == ["rank_0.pt"]
        assert res["segment"].endswith("ckpt_shm_0")
    # the two windows tile the segment at a 2 MiB boundary
    (a0, a1), (b0, b1) = r0["window"], r1["window"]
    assert a0 == 0 and a1 == b0 and b1 == r0["total"] and a1 % (2 << 20) == 0 and a1 > 0
    # meta plane: no SharedDict.set at all, the control segment carries the saves
    assert r0["dict_sets"] == 0 and r0["ctl_publishes"] == 4   # 2 saves x (announce, finish)
    assert r1["dict_sets"] == 0 and r1["ctl_publishes"] == 0   # followers publish nothing


@pytest.mark.timeout(300)
def test_only_rank0_calling_save_does_not_hang():
    """The cooperative default must not turn a rank-0-only save loop into a deadlock: the
    leader notices that nobody joins, saves the whole image alone and stops waiting."""
    r0, r1 = _run("rank0_only")
    assert r0["load_ok"] and r1["load_ok"]
    assert r0["coop_wanted_after"] is False
    assert 0.4 < r0["first_s"] < 10 and r0["second_s"] < r0["first_s"]


@pytest.mark.timeout(300)
def test_replicated_state_only_local_rank0_saves():
    r0, r1 = _run("replicated")
    assert r0["saving_ranks"] == [0] and r1["saving_ranks"] == [0]
    assert r0["cached_step"] == 7 and r1["cached_step"] == -1
    assert r0["files"] == ["rank_0.pt"] and r1["files"] == ["rank_0.pt"]
    assert r0["load_ok"] and r1["load_ok"]
