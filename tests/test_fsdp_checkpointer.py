"""FsdpShardCheckpointer / FsdpFullCheckpointer with a real FSDP-wrapped module
(world_size 1): save to memory, restore into a fresh model+optimizer, identical
forward logits (reference: fsdp_ckpt_test.py test_fsdp_checkpointer).  GPU only:
torch 2.11's FSDP refuses to wrap without an accelerator; the CPU coverage of the
same engine is tests/test_fsdp_engine.py."""

import os

import pytest
import torch
import torch.distributed as dist
import torch.nn as nn
from torch.distributed.fsdp import FullyShardedDataParallel as FSDP

from dlrover_b200.ckpt_saver import AsyncCheckpointSaver
from dlrover_b200.flash_checkpoint.api import StorageType
from dlrover_b200.flash_checkpoint.fsdp import FsdpFullCheckpointer, FsdpShardCheckpointer


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(64, 128)
        self.b = nn.Linear(128, 32)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def _setup(device, monkeypatch, backend):
    monkeypatch.setenv("MASTER_ADDR", "127.0.0.1")
    monkeypatch.setenv("MASTER_PORT", str(29700 + os.getpid() % 1500))
    for k, v in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0"),
                 ("LOCAL_WORLD_SIZE", "1")):
        monkeypatch.setenv(k, v)
    dist.init_process_group(backend, rank=0, world_size=1)
    AsyncCheckpointSaver.start_async_saving_ckpt()


def _train_one_step(device, seed):
    torch.manual_seed(seed)
    model = FSDP(Net().to(device), device_id=device if device.type == "cuda" else None)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
    x = torch.randn(8, 64, device=device)
    model(x).sum().backward()
    opt.step()
    opt.zero_grad()
    return model, opt


def _roundtrip(ckpt_cls, tmp_path, device, storage_type):
    ckpt = ckpt_cls(str(tmp_path))
    model, opt = _train_one_step(device, 0)
    x = torch.randn(4, 64, device=device)
    with torch.no_grad():
        want = model(x).clone()
    ckpt.save_checkpoint(10, model, opt, {"step": 10}, storage_type=storage_type)
    if storage_type == StorageType.DISK:
        ckpt.wait_latest_checkpoint(timeout=120)
        assert (tmp_path / "dlrover_latest.txt").read_text() == "10"
    else:
        ckpt.wait_memory_save(60)
    model2, opt2 = _train_one_step(device, 1)  # different weights
    with torch.no_grad():
        assert not torch.equal(model2(x), want)
    extra = ckpt.load_checkpoint(model2, opt2)
    # the sharded checkpointer restores the keys it asks DCP for: model, optim and
    # "step" (reference fsdp.py:122-127); the full one returns every extra key
    assert extra.get("step") == 10
    with torch.no_grad():
        assert torch.equal(model2(x), want)  # bit-identical logits
    ckpt.engine.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cls", [FsdpShardCheckpointer, FsdpFullCheckpointer])
@pytest.mark.parametrize("storage_type", [StorageType.MEMORY, StorageType.DISK])
def test_fsdp_checkpointers_cuda(cuda_device, run_env, monkeypatch, tmp_path, cls, storage_type):
    _setup(cuda_device, monkeypatch, "nccl")
    try:
        _roundtrip(cls, tmp_path, cuda_device, storage_type)
    finally:
        dist.destroy_process_group()
