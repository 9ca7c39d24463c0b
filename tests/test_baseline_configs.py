"""BASELINE.json configs[3] and [4] as parity cases (scaled): the dict shapes the
frameworks hand to the engines — ZeRO-3 style flat fp32 partitions through
DeepSpeedCheckpointEngine, Megatron TPxPP model shard + distributed-optimizer
shard through MegatronDistCheckpointEngine — compared byte-for-byte with the
oracle image."""

import numpy as np
import pytest
import torch

from dlrover_b200.ckpt_saver import AsyncCheckpointSaver, DLROVER_CKPT_CONFIG_KEY
from dlrover_b200.common.constants import CheckpointConstant
from dlrover_b200.common.storage import PosixDiskStorage
from dlrover_b200.flash_checkpoint.engine import (
    DeepSpeedCheckpointEngine,
    MegatronDistCheckpointEngine,
)
from dlrover_b200 import shapes
from oracle import shm_layout as oracle

MODEL = CheckpointConstant.MODEL_STATES_NAME
OPTIM = CheckpointConstant.OPTIM_STATES_NAME


@pytest.fixture
def agent(run_env):
    AsyncCheckpointSaver.start_async_saving_ckpt()
    yield


def _image_equals_oracle(engine, state):
    seg = np.frombuffer(engine._shm_handler.shared_memory.buf, dtype=np.uint8)
    meta = engine._shm_handler.metadata.get()
    conf = meta[DLROVER_CKPT_CONFIG_KEY]
    assert conf.writing_shm is False
    _, want = oracle.serialize(state)
    assert seg.size == want.size and np.array_equal(seg, want)


def _zero3_state(device, numel):
    """What DeepSpeed ZeRO-3 saves per rank: a few huge flat fp32 partitions
    (70B/8 ranks -> 3 x 35 GB; here `numel` elements each) + small bookkeeping."""
    def flat(seed):
        return shapes.fill_(torch.empty(numel, dtype=torch.float32, device=device), seed)

    optim = {
        "optimizer_state_dict": {
            "fp32_flat_groups": [flat(1)],
            "optimizer_state_dict": {
                "state": {0: {"step": torch.tensor(1000.0), "exp_avg": flat(2),
                              "exp_avg_sq": flat(3)}},
                "param_groups": [{"lr": 1e-5, "betas": (0.9, 0.95), "params": [0]}],
            },
            "zero_stage": 3, "partition_count": [8], "ds_version": "0.14.0",
        },
        "ds_config": {"zero_optimization": {"stage": 3}},
    }
    model = {"module": None, "buffer_names": [], "param_shapes": [{"w": (4096, 4096)}],
             "global_steps": 1000, "dp_world_size": 8, "mp_world_size": 1}
    return {MODEL: model, OPTIM: optim}


def _roundtrip_zero3(tmp_path, device, numel, async_drain):
    engine = DeepSpeedCheckpointEngine(str(tmp_path), PosixDiskStorage(), global_shard_num=1,
                                       zero_stage=3, async_drain=async_drain)
    state = _zero3_state(device, numel)
    paths = {MODEL: str(tmp_path / "1000" / "zero_pp_rank_0_mp_rank_00_model_states.pt"),
             OPTIM: str(tmp_path / "1000" / "zero_pp_rank_0_mp_rank_00_optim_states.pt")}
    assert engine.save_to_memory(1000, dict(state), paths)
    assert engine.wait_memory_save(120)
    _image_equals_oracle(engine, {**state, DLROVER_CKPT_CONFIG_KEY: None})
    loaded = engine.load()
    got = loaded[OPTIM]["optimizer_state_dict"]["fp32_flat_groups"][0]
    assert torch.equal(got, state[OPTIM]["optimizer_state_dict"]["fp32_flat_groups"][0].cpu())
    assert loaded[MODEL]["global_steps"] == 1000
    del loaded, got
    engine.close()


def test_zero3_flat_partitions_cpu(agent, tmp_path):
    _roundtrip_zero3(tmp_path, "cpu", 100_003, async_drain=False)


@pytest.mark.gpu
def test_zero3_flat_partitions_cuda(cuda_device, agent, tmp_path):
    # 3 x 256 MiB fp32 (+ a 4-byte CPU step scalar that misaligns what follows)
    _roundtrip_zero3(tmp_path, "cuda", (64 << 20) + 3, async_drain=True)


def _megatron_state(device, scale):
    """(tp,pp) model shard of a Mixtral-like layer stack + this rank's
    distributed-optimizer shard (bucket -> group -> order -> tensors)."""
    h, ffn, experts = int(4096 * scale), int(14336 * scale), 8
    model = {"args": {"tp": 2, "pp": 2}, "iteration": 20, "checkpoint_version": 3.0,
             "model": {}}
    for l in range(2):
        p = f"decoder.layers.{l}."
        model["model"][p + "self_attention.linear_qkv.weight"] = shapes.fill_(
            torch.empty((3 * h // 2, h), dtype=torch.bfloat16, device=device), l)
        model["model"][p + "mlp.router.weight"] = shapes.fill_(
            torch.empty((experts, h), dtype=torch.bfloat16, device=device), 10 + l)
        for e in range(experts // 2):
            model["model"][p + f"mlp.experts.local_experts.{e}.linear_fc1.weight"] = shapes.fill_(
                torch.empty((ffn, h), dtype=torch.bfloat16, device=device), 20 + e)
    optim = {0: {0: {}}}
    for i, (k, t) in enumerate(list(model["model"].items())[:6]):
        n = t.numel() // 4  # this DP rank's slice
        optim[0][0][i] = {
            "param": shapes.fill_(torch.empty(n, dtype=torch.float32, device=device), 100 + i),
            "exp_avg": shapes.fill_(torch.empty(n, dtype=torch.float32, device=device), 200 + i),
            "exp_avg_sq": shapes.fill_(torch.empty(n, dtype=torch.float32, device=device), 300 + i),
        }
    return {MODEL: model, OPTIM: optim}


def _roundtrip_megatron(tmp_path, device, scale, async_drain):
    engine = MegatronDistCheckpointEngine(str(tmp_path), PosixDiskStorage(),
                                          async_drain=async_drain)
    state = _megatron_state(device, scale)
    paths = {MODEL: str(tmp_path / "iter_0000020" / "mp_rank_00_000" / "model_optim_rng.pt"),
             OPTIM: str(tmp_path / "iter_0000020" / "rank_00000" / "distrib_optim.pt")}
    assert engine.save_to_memory(20, dict(state), paths)
    assert engine.wait_memory_save(120)
    _image_equals_oracle(engine, {**state, DLROVER_CKPT_CONFIG_KEY: None})
    step, loaded = engine.load()
    assert step == 20 and loaded[MODEL]["iteration"] == 20
    k = next(iter(state[MODEL]["model"]))
    assert torch.equal(loaded[MODEL]["model"][k], state[MODEL]["model"][k].cpu())
    del loaded
    engine.close()


def test_megatron_shards_cpu(agent, tmp_path):
    _roundtrip_megatron(tmp_path, "cpu", 1 / 64, async_drain=False)


@pytest.mark.gpu
def test_megatron_shards_cuda(cuda_device, agent, tmp_path):
    _roundtrip_megatron(tmp_path, "cuda", 1 / 8, async_drain=True)
