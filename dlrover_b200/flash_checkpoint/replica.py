"""Cross-node in-memory replicas of checkpoint shards (multi-node jobs only).

Interface and group arithmetic follow the reference @ 468d632,
dlrover/trainer/torch/flash_checkpoint/replica.py: CkptReplicaManger (:28-70),
ShardCkptReplicaManager (:73-244: peers = same local rank on the
`replica_count` nodes of a backup group; after every save the group all-gathers
the raw segment bytes over gloo into extra segments), FullCkptReplicaManager
(:247-352: replicated checkpoints need no backup; a replaced node receives a
broadcast from any node that still holds the bytes).

On a single NVSwitch node there is nothing to replicate to; this module keeps
the reference's host-side (gloo) byte exchange because the payload already
lives in host shared memory when backup() runs.
"""

from __future__ import annotations

from abc import ABCMeta, abstractmethod
from typing import Dict, List

import torch
import torch.distributed as dist

from ..common import env_utils
from ..shm_handler import DLROVER_CKPT_CONFIG_KEY, CheckpointConfig, SharedMemoryHandler


def _segment_bytes(handler: SharedMemoryHandler) -> torch.Tensor:
    """uint8 COPY of the handler's segment (empty tensor if not mapped)."""
    if handler.shared_memory is None:
        return torch.empty(0, dtype=torch.uint8)
    return torch.frombuffer(handler.shared_memory.buf, dtype=torch.uint8).clone()


class CkptReplicaManger(metaclass=ABCMeta):
    def __init__(self, replica_count) -> None:
        self.replica_count = replica_count
        self.local_rank = env_utils.get_local_rank()
        self.local_world_size = env_utils.get_local_world_size()
        self.node_rank = env_utils.get_node_rank()
        self.node_num = env_utils.get_node_num()
        self.current_device = torch.device("cpu")
        self.rank = dist.get_rank() if dist.is_initialized() else env_utils.get_rank()
        self.backup_ranks: List[int] = []
        self._rank_shms: Dict[int, SharedMemoryHandler] = {}
        self._backup_group = None

    @staticmethod
    def create_replica_manager(shard_num, replica_count):
        if shard_num == 1:
            return FullCkptReplicaManager(replica_count)
        return ShardCkptReplicaManager(replica_count)

    def has_replica(self):
        return self.replica_count > 0

    def _make_group(self):
        if dist.is_initialized() and self.replica_count > 0:
            self._backup_group = dist.new_group(backend="gloo", ranks=self.backup_ranks)

    def _max_size_in_group(self, local_numel: int) -> int:
        mine = torch.tensor([local_numel], dtype=torch.long)
        sizes = [torch.zeros(1, dtype=torch.long) for _ in self.backup_ranks]
        dist.all_gather(sizes, mine, group=self._backup_group)
        return int(max(int(s.item()) for s in sizes))

    def _padded(self, raw: torch.Tensor) -> torch.Tensor:
        size = self._max_size_in_group(raw.numel())
        if raw.numel() == size:
            return raw
        out = torch.zeros(size, dtype=torch.uint8)
        out[: raw.numel()] = raw
        return out

    @abstractmethod
    def backup(self, shm_handler: SharedMemoryHandler):
        """Exchange this rank's shard with its backup group."""

    @abstractmethod
    def gather(self, shm_handler: SharedMemoryHandler):
        """Fetch this rank's shard back from the group: (bytes, meta) or (None, None/{})"""


class ShardCkptReplicaManager(CkptReplicaManger):
    """Sharded checkpoints: rank r's peers are the ranks with the same local
    rank on the other nodes of its backup group, e.g. replica_count=2,
    8 ranks/node -> local rank 0 of nodes {0,1} = ranks [0, 8]."""

    def __init__(self, replica_count=0) -> None:
        super().__init__(replica_count)
        self.backup_ranks = self._get_backup_ranks(replica_count)
        self._make_group()

    def _get_backup_ranks(self, replica_count):
        if replica_count <= 0:
            return []
        first_node = (self.node_rank // replica_count) * replica_count
        return [(first_node + i) * self.local_world_size + self.local_rank
                for i in range(replica_count)]

    def _exchange(self, raw: torch.Tensor, meta):
        """all-gather (padded) segment bytes and meta dicts inside the group."""
        mine = self._padded(raw)
        world = dist.get_world_size(group=self._backup_group)
        blobs = [torch.empty(mine.numel(), dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(blobs, mine, group=self._backup_group)
        metas = [None] * world
        dist.all_gather_object(metas, meta, group=self._backup_group)
        return blobs, metas

    def backup(self, shm_handler: SharedMemoryHandler):
        if self.replica_count == 0:
            return
        assert shm_handler.shared_memory is not None
        blobs, metas = self._exchange(_segment_bytes(shm_handler), shm_handler.metadata.get())
        self._rank_shms[self.rank] = shm_handler
        for blob, meta in zip(blobs, metas):
            owner: CheckpointConfig = meta[DLROVER_CKPT_CONFIG_KEY]
            if owner.rank == self.rank:
                continue
            peer = self._rank_shms.get(owner.rank)
            if peer is None:
                peer = SharedMemoryHandler(local_rank=owner.rank)
                peer.init_shared_memory(create=True, size=blob.numel())
                self._rank_shms[owner.rank] = peer
            torch.frombuffer(peer.shared_memory.buf, dtype=torch.uint8).copy_(blob)
            peer.metadata.set(meta)

    def gather(self, shm_handler: SharedMemoryHandler):
        """One exchange round per group member; in the round of rank r every
        member offers what it holds for r and r keeps the first non-empty one."""
        holders: Dict[int, SharedMemoryHandler] = {}
        for rank in self.backup_ranks:
            if rank == self.rank:
                holders[rank] = shm_handler
            elif rank in self._rank_shms:
                # the copy (and its meta dict) this process received in backup()
                holders[rank] = self._rank_shms[rank]
            else:
                h = SharedMemoryHandler(local_rank=rank)
                h.init_shared_memory()
                holders[rank] = h
        return self._gather_owner_checkpoint(holders)

    def _gather_owner_checkpoint(self, shm_handlers):
        """One exchange round per group member.  `shm_handlers`: what this rank
        holds for every member — a dict keyed by rank, or a sequence in
        backup_ranks order."""
        found_bytes, found_meta = None, {}
        for i, rank in enumerate(self.backup_ranks):
            h = shm_handlers[rank] if isinstance(shm_handlers, dict) else shm_handlers[i]
            if h.shared_memory:
                raw, meta = _segment_bytes(h), h.metadata.get()
            else:
                raw, meta = torch.empty(0, dtype=torch.uint8), {}
            blobs, metas = self._exchange(raw, meta)
            if rank != self.rank:
                continue
            for blob, m in zip(blobs, metas):
                if m:
                    found_bytes, found_meta = blob, m
                    break
        return found_bytes, found_meta


class FullCkptReplicaManager(CkptReplicaManger):
    """Replicated checkpoints: every node already holds a full copy, so backup
    is a no-op; gather broadcasts from the first node that still has it."""

    def __init__(self, replica_count=0) -> None:
        super().__init__(replica_count)
        self.backup_ranks = self._get_backup_ranks()
        self._make_group()

    def _get_backup_ranks(self):
        """Local rank 0 of every node."""
        return [node * self.local_world_size for node in range(self.node_num)]

    def backup(self, shm_handler: SharedMemoryHandler):
        pass

    def gather(self, shm_handler: SharedMemoryHandler):
        if self.rank not in self.backup_ranks:
            return None, None
        have = torch.tensor([1 if shm_handler.shared_memory else 0], dtype=torch.int8)
        flags = [torch.zeros(1, dtype=torch.int8) for _ in self.backup_ranks]
        dist.all_gather(flags, have, group=self._backup_group)
        owners = [r for r, f in zip(self.backup_ranks, flags) if int(f.item()) == 1]
        if not owners:
            return None, None
        payload = self._padded(_segment_bytes(shm_handler))
        dist.broadcast(payload, src=owners[0], group=self._backup_group)
        metas = [shm_handler.metadata.get()]
        dist.broadcast_object_list(metas, src=owners[0], group=self._backup_group)
        if payload.numel() == 0:
            return None, None
        return payload, metas[0]
