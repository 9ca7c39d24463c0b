"""DeepSpeed / Megatron-LM adapters, exercised with tiny stand-ins for the
frameworks (neither is installed here; the reference's tests use the same
approach: deepspeed_ckpt_test.py:40-77, megatron_ckpt_test.py:93-124)."""

import argparse
import os
import time
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

from dlrover_b200.ckpt_saver import AsyncCheckpointSaver
from dlrover_b200.common.constants import CheckpointConstant
from dlrover_b200.flash_checkpoint import megatron as fc_megatron
from dlrover_b200.flash_checkpoint import megatron_dist_ckpt as fc_dist
from dlrover_b200.flash_checkpoint.api import StorageType
from dlrover_b200.flash_checkpoint.deepspeed import AsyncCheckpointAgent, DeepSpeedCheckpointer
from dlrover_b200.flash_checkpoint.engine import MegatronDistCheckpointEngine
from dlrover_b200.common.storage import PosixDiskStorage
from tests.util import tree_equal

MODEL = CheckpointConstant.MODEL_STATES_NAME
OPTIM = CheckpointConstant.OPTIM_STATES_NAME


class SimpleNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(64, 32)
        self.fc2 = nn.Linear(32, 10)


@pytest.fixture
def agent(run_env):
    AsyncCheckpointSaver.start_async_saving_ckpt()
    yield
    for cls in (fc_megatron.MegatronCheckpointer, fc_dist.MegatronDistCheckpointer):
        if "_singleton_obj" in cls.__dict__:
            try:
                cls._singleton_obj.engine.close()
            except Exception:
                pass
            delattr(cls, "_singleton_obj")


class FakeDeepSpeedEngine:
    def __init__(self, model, optimizer):
        self.model, self.optimizer = model, optimizer
        self.save_non_zero_checkpoint = False
        self.global_rank = 0

    def zero_optimization(self):
        return False

    def zero_optimization_stage(self):
        return 0

    def save_checkpoint(self, save_dir, tag, client_state, save_latest):
        os.makedirs(os.path.join(save_dir, str(tag)), exist_ok=True)
        torch.save(self.model.state_dict(), os.path.join(save_dir, str(tag), "model_states.pt"))
        torch.save(self.optimizer.state_dict(),
                   os.path.join(save_dir, str(tag), "optim_states.pt"))
        with open(os.path.join(save_dir, "latest"), "w") as f:
            f.write(str(tag))

    def load_checkpoint(self, load_dir, tag, load_module_strict, load_optimizer_states,
                        load_lr_scheduler_states, load_module_only, custom_load_fn):
        self.model.load_state_dict(torch.load(os.path.join(load_dir, str(tag),
                                                           "model_states.pt")))
        self.optimizer.load_state_dict(torch.load(os.path.join(load_dir, str(tag),
                                                               "optim_states.pt")))
        return os.path.join(load_dir, str(tag)), {}


def test_deepspeed_checkpointer(agent, tmp_path):
    """deepspeed_ckpt_test.py:100-160: 9640-byte segment, file names, `latest`."""
    model = SimpleNet()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.001)
    engine = FakeDeepSpeedEngine(model, opt)
    d = str(tmp_path)
    ckpt = DeepSpeedCheckpointer(engine, d, async_drain=False)
    assert engine.save_non_zero_checkpoint is True  # below ZeRO-3, local rank 0
    before = torch.save
    ckpt.save_checkpoint(d, 100, storage_type=StorageType.MEMORY)
    assert torch.save is before
    assert ckpt._async_save_engine._shm_handler._buffer_size == 9640
    # memory-only: DeepSpeed's tag dir and `latest` are rolled back
    assert not os.path.exists(tmp_path / "100") and not os.path.exists(tmp_path / "latest")
    want = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        for p in model.parameters():
            p.zero_()
    ckpt.load_checkpoint(d, 100)  # served from shared memory, no files exist
    assert tree_equal(dict(model.state_dict()), want)
    ckpt.save_checkpoint(d, 200, storage_type=StorageType.DISK)
    ckpt.wait_latest_checkpoint(timeout=60)
    assert sorted(os.listdir(tmp_path / "200")) == ["model_states.pt", "optim_states.pt"]
    assert (tmp_path / "latest").read_text() == "200"
    assert (tmp_path / "dlrover_latest.txt").read_text() == "200"
    with pytest.raises(ValueError):
        ckpt.save_checkpoint(d, 300, storage_type="x")
    ckpt._async_save_engine.close()


def test_async_checkpoint_agent_naming(tmp_path):
    agent_ = AsyncCheckpointAgent(PosixDiskStorage())
    agent_.save({"a": 1}, "/x/mp_rank_00_model_states.pt")
    agent_.save({"b": 2}, "/x/zero_pp_rank_0_mp_rank_00_optim_states.pt")
    agent_.save({"c": 3}, "/x/other.pt")
    assert set(agent_.state_dict) == {MODEL, OPTIM, "other.pt"}
    assert agent_.load("/y/model_states.pt") == {"a": 1}
    p = str(tmp_path / "f.pt")
    torch.save({"z": 9}, p)
    assert agent_.load(p) == {"z": 9}
    with open(tmp_path / "stream.pt", "wb") as f:
        agent_.save({"q": 1}, f)  # non-str path goes to the real torch.save
    assert torch.load(tmp_path / "stream.pt") == {"q": 1}


def test_megatron_save_load(agent, tmp_path, monkeypatch):
    """megatron_ckpt_test.py:83-164."""
    model = SimpleNet()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.001)
    d = str(tmp_path)
    suffix = "model_optim_rng.pt"

    def get_args():
        ns = argparse.Namespace()
        ns.save = d
        return ns

    def fake_save(iteration, model, optimizer, sched):
        path = os.path.join(d, "iter_{:07d}".format(iteration), "mp_rank_00", suffix)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({"iteration": iteration, "model": model.state_dict(),
                    "optimizer": optimizer.state_dict()}, path)
        with open(os.path.join(d, "latest_checkpointed_iteration.txt"), "w") as f:
            f.write(str(iteration))

    def fake_load(model, optimizer, sched, load_arg="load", strict=True):
        path = os.path.join(d, "iter_{:07d}".format(20), "mp_rank_00", suffix)
        sd = torch.load(path)
        model.load_state_dict(sd["model"])
        optimizer.load_state_dict(sd["optimizer"])
        return sd["iteration"]

    monkeypatch.setattr(fc_megatron, "megatron_save", fake_save)
    monkeypatch.setattr(fc_megatron, "megatron_load", fake_load)
    monkeypatch.setattr(fc_megatron, "get_args", get_args)
    monkeypatch.setenv("DLROVER_B200_ASYNC_DRAIN", "0")

    fc_megatron.save_checkpoint(10, model, opt, None, storage_type=StorageType.MEMORY)
    saver = fc_megatron.MegatronCheckpointer.singleton_instance(d)
    assert saver.engine._shm_handler._buffer_size == 9640  # megatron_ckpt_test.py:133-135
    # memory-only: Megatron's directory/tracker rolled back
    assert not os.path.exists(tmp_path / "iter_0000010")
    assert not os.path.exists(tmp_path / "latest_checkpointed_iteration.txt")
    fc_megatron.save_checkpoint(20, model, opt, None, storage_type=StorageType.DISK)
    fc_megatron.wait_latest_checkpoint(timeout=60)
    assert (tmp_path / "latest_checkpointed_iteration.txt").read_text() == "20"
    assert (tmp_path / "dlrover_latest.txt").read_text() == "20"
    assert os.path.exists(tmp_path / "iter_0000020" / "mp_rank_00" / suffix)
    want = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        for p in model.parameters():
            p.zero_()
    assert fc_megatron.load_checkpoint(model, opt, None) == 20
    assert tree_equal(dict(model.state_dict()), want)
    with pytest.raises(ValueError):
        saver.save({}, "/x/unknown_name.pt")
    with pytest.raises(ValueError):
        fc_megatron.save_checkpoint(30, model, opt, None, storage_type="bogus")


class FakeDistOptimizer:
    """Shape of Megatron's DistributedOptimizer that the shard walkers touch."""

    def __init__(self, params, device):
        self.optimizer = SimpleNamespace(param_groups=[{"params": []}], state={})
        self.model_param_group_index_map = {}
        param_map = {}
        for order, p in enumerate(params):
            main = p.detach().clone().float().to(device)
            self.optimizer.param_groups[0]["params"].append(main)
            self.optimizer.state[main] = {"exp_avg": torch.full_like(main, 0.5 + order),
                                          "exp_avg_sq": torch.full_like(main, 2.0 + order)}
            self.model_param_group_index_map[p] = (0, order)
            param_map[p] = {}
        self.gbuf_ranges = [{torch.float32: [{"param_map": param_map}]}]

    def get_parameter_state(self):  # marker attribute for chained optimizers
        return fc_dist.get_parameter_state(self)

    load_parameter_state_from_state_dict = True


def _dist_roundtrip(tmp_path, device):
    model = SimpleNet().to(device)
    dopt = FakeDistOptimizer(list(model.parameters()), device)
    state = fc_dist.get_parameter_state(dopt)
    assert sorted(state[0][0]) == [0, 1, 2, 3]
    assert set(state[0][0][0]) == {"param", "exp_avg", "exp_avg_sq"}
    assert fc_dist.get_dist_optimizer_checkpoint_name("/tmp", 100) == \
        "/tmp/iter_0000100/rank_00000/distrib_optim.pt"  # megatron_ckpt_test.py:198-199
    engine = MegatronDistCheckpointEngine(str(tmp_path), PosixDiskStorage(), async_drain=False)
    path = fc_dist.get_dist_optimizer_checkpoint_name(str(tmp_path), 100)
    assert engine.save_to_memory(100, {OPTIM: state}, {OPTIM: path})
    want = {o: {k: v.clone() for k, v in t.items()} for o, t in state[0][0].items()}
    for t in state[0][0].values():
        for v in t.values():
            v.zero_()
    # per-tensor path (what the storage fallback uses)
    step, loaded = engine.load()
    assert step == 100
    fc_dist.load_parameter_state_from_state_dict(dopt, loaded[OPTIM])
    assert tree_equal(state[0][0], want)
    for t in state[0][0].values():
        for v in t.values():
            v.zero_()
    # one-shot scatter path
    fc_dist.bind_megatron(SimpleNamespace(ChainedOptimizer=type("Chained", (), {})))
    try:
        assert fc_dist._restore_shards_from_memory(engine, dopt) is True
    finally:
        fc_dist.bind_megatron(None)
    assert tree_equal(state[0][0], want)
    engine.close()


def test_dist_optimizer_shards_cpu(agent, tmp_path):
    _dist_roundtrip(tmp_path, "cpu")


@pytest.mark.gpu
def test_dist_optimizer_shards_cuda(cuda_device, agent, tmp_path):
    _dist_roundtrip(tmp_path, "cuda")


def test_megatron_deletion_strategies(tmp_path):
    for step in (100, 200, 300):
        os.makedirs(tmp_path / "iter_{:07d}".format(step))
    keep = fc_dist.KeepLatestStepStrategy(2, str(tmp_path))
    import shutil
    for step in (100, 200, 300):
        keep.clean_up(step, shutil.rmtree)
    assert sorted(os.listdir(tmp_path)) == ["iter_0000300"] or \
        sorted(os.listdir(tmp_path)) == ["iter_0000200", "iter_0000300"]
    interval = fc_dist.KeepStepIntervalStrategy(200, str(tmp_path))
    os.makedirs(tmp_path / "iter_0000100", exist_ok=True)
    interval.clean_up(100, shutil.rmtree)
    interval.clean_up(200, shutil.rmtree)
    assert not os.path.exists(tmp_path / "iter_0000100")
    assert isinstance(fc_dist.get_checkpoint_storage(keep), fc_dist.PosixStorageWithDeletion)


def test_hf_checkpointer_records_and_persists(agent, tmp_path):
    """The part of the HF adapter that does not need a live Trainer: entries
    recorded from torch.save / safetensors are persisted under their paths."""
    from dlrover_b200.flash_checkpoint.hf_trainer import HfDdpCheckpointer, _SafetensorsRecorder

    ckpt = HfDdpCheckpointer(str(tmp_path))
    agent_ = ckpt.ckpt_agent
    agent_.safetensors_metadata = {}
    out = tmp_path / "checkpoint-2"
    os.makedirs(out)
    model = SimpleNet()
    _SafetensorsRecorder(agent_)(dict(model.state_dict()), str(out / "model.safetensors"),
                                 metadata={"format": "pt"})
    agent_.save({"state": {}, "param_groups": [{"lr": 0.1}]}, str(out / "optimizer.pt"))
    agent_.save({"python": 1}, str(out / "rng_state.pth"))
    assert ckpt.save_checkpoint_to_storage(2)
    ckpt.async_save_engine.wait_latest_checkpoint(timeout=60)
    assert sorted(os.listdir(out)) == ["model.safetensors", "optimizer.pt", "rng_state.pth"]
    from safetensors.torch import load_file

    back = load_file(str(out / "model.safetensors"))
    want = model.state_dict()
    assert set(back) == set(want) and all(torch.equal(back[k], want[k]) for k in want)
    assert torch.load(out / "optimizer.pt", weights_only=False)["param_groups"][0]["lr"] == 0.1
    ckpt.async_save_engine.close()
