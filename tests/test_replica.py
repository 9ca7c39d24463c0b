"""Replica managers: group arithmetic (known answers from the reference's
checkpoint_backup_test.py:113-127) and a 2-process gloo backup/gather."""

import os
import sys
import tempfile
import uuid

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dlrover_b200.flash_checkpoint.replica import (
    CkptReplicaManger,
    FullCkptReplicaManager,
    ShardCkptReplicaManager,
)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_backup_rank_arithmetic(monkeypatch):
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    monkeypatch.setenv("NODE_RANK", "1")
    monkeypatch.setenv("NODE_NUM", "4")
    shard = ShardCkptReplicaManager(replica_count=0)
    assert shard._get_backup_ranks(2) == [0, 8]  # nodes {0,1}, local rank 0
    assert shard._get_backup_ranks(0) == [] and not shard.has_replica()
    monkeypatch.setenv("NODE_RANK", "3")
    monkeypatch.setenv("LOCAL_RANK", "5")
    assert ShardCkptReplicaManager(0)._get_backup_ranks(2) == [21, 29]  # nodes {2,3}
    full = FullCkptReplicaManager(replica_count=0)
    assert full.backup_ranks == [0, 8, 16, 24]
    assert isinstance(CkptReplicaManger.create_replica_manager(1, 0), FullCkptReplicaManager)
    assert isinstance(CkptReplicaManger.create_replica_manager(4, 0), ShardCkptReplicaManager)


def _worker(rank, world, port, run_id, out_dir):
    sys.path.insert(0, ROOT)
    # two "nodes" with one rank each
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": "0", "WORLD_SIZE": str(world),
                       "LOCAL_WORLD_SIZE": "1", "NODE_RANK": str(rank), "NODE_NUM": str(world),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port),
                       "TORCHELASTIC_RUN_ID": f"{run_id}n{rank}", "ROLE_NAME": "dlrover-trainer",
                       "DLROVER_LOG_LEVEL": "ERROR"})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dlrover_b200.flash_checkpoint.replica import ShardCkptReplicaManager
    from dlrover_b200.shm_handler import (DLROVER_CKPT_CONFIG_KEY, CheckpointConfig,
                                          SharedMemoryHandler)

    result = {}
    try:
        mgr = ShardCkptReplicaManager(replica_count=2)
        assert mgr.backup_ranks == [0, 1]
        handler = SharedMemoryHandler(0, host=True)
        n = 1632 if rank == 0 else 800  # different shard sizes: padding path
        sd = {"w": torch.full((n // 4,), float(rank + 1)),
              DLROVER_CKPT_CONFIG_KEY: CheckpointConfig(rank=rank, step=5, paths={})}
        handler.save_state_dict(sd)
        mgr.backup(handler)
        peer = 1 - rank
        held = mgr._rank_shms[peer]
        peer_sd = held.load_state_dict()
        result["peer_value"] = float(peer_sd["w"][0])
        result["peer_numel"] = int(peer_sd["w"].numel())
        del peer_sd
        # rank 1 "loses its node": fresh handler with nothing, gets the shard back
        dist.barrier()
        if rank == 1:
            handler.shared_memory.unlink()
            handler.shared_memory.close()
            handler.shared_memory = None
            handler.metadata.set({})
        blob, meta = mgr.gather(handler)
        if rank == 1:
            result["recovered"] = bool(meta) and int(blob.numel()) >= 800 and \
                meta[DLROVER_CKPT_CONFIG_KEY].rank == 1
            w = torch.frombuffer(bytes(blob[:800].numpy().tobytes()), dtype=torch.float32)
            result["recovered_value"] = float(w[0])
        dist.barrier()
        for h in list(mgr._rank_shms.values()) + [handler]:
            try:
                h.unlink()
                h.close()
            except Exception:
                pass
    finally:
        torch.save(result, os.path.join(out_dir, f"r{rank}.pt"))
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_shard_backup_and_gather_two_nodes():
    run_id = "rp" + uuid.uuid4().hex[:8]
    port = 29300 + os.getpid() % 500
    with tempfile.TemporaryDirectory() as out_dir:
        mp.spawn(_worker, args=(2, port, run_id, out_dir), nprocs=2, join=True)
        r0 = torch.load(os.path.join(out_dir, "r0.pt"))
        r1 = torch.load(os.path.join(out_dir, "r1.pt"))
    assert r0["peer_value"] == 2.0 and r0["peer_numel"] == 200
    assert r1["peer_value"] == 1.0 and r1["peer_numel"] == 408
    assert r1["recovered"] and r1["recovered_value"] == 2.0
    import glob
    import shutil
    for r in (0, 1):
        shutil.rmtree(os.path.join("/tmp/ckpt_sock", f"{run_id}n{r}"), ignore_errors=True)
        for f in glob.glob(f"/dev/shm/{run_id}n{r}_*"):
            os.unlink(f)
