"""Flash checkpointer for DeepSpeed engines.

Reference @ 468d632: dlrover/trainer/torch/flash_checkpoint/deepspeed.py —
AsyncCheckpointAgent (:45-95) records what DeepSpeedEngine.save_checkpoint
would torch.save ("*model_states.pt" -> "model_states", "*optim_states.pt" ->
"optim_states", anything else under its file name) together with the path;
DeepSpeedCheckpointer (:98-264) runs the engine's own save/load with
torch.save/torch.load swapped for that recorder and hands the captured dicts to
DeepSpeedCheckpointEngine.  Under ZeRO-3 the captured tensors are a handful of
very large flat fp32 partitions per rank — one descriptor each for the gather
kernel.

deepspeed itself is imported lazily (it is only needed for the ZeRO stage enum
and is absent from many environments, this one included).
"""

from __future__ import annotations

import os
from typing import Dict

import torch.distributed as dist

from ..ckpt_saver import DeepSpeedCheckpointSaver
from ..common import env_utils
from ..common.constants import CheckpointConstant
from ..common.storage import CheckpointStorage, get_checkpoint_storage
from .api import Checkpointer, StorageType
from .engine import DeepSpeedCheckpointEngine
from .torch_io_patch import (
    patched_torch_load,
    patched_torch_save,
    torch_native_load,
    torch_native_save,
)

_DS_MODEL_SD_FILE_SUFFIX = "model_states.pt"
_DS_OPTIM_SD_FILE_SUFFIX = "optim_states.pt"
_ZERO_STAGE_WEIGHTS = 3  # deepspeed.runtime.zero.config.ZeroStageEnum.weights


def _state_name(path: str, strict: bool = False) -> str:
    if path.endswith(_DS_MODEL_SD_FILE_SUFFIX):
        return CheckpointConstant.MODEL_STATES_NAME
    if path.endswith(_DS_OPTIM_SD_FILE_SUFFIX):
        return CheckpointConstant.OPTIM_STATES_NAME
    return "" if strict else path.split("/")[-1]


class AsyncCheckpointAgent:
    """Stands in for torch.save/torch.load (and for DeepSpeed's pluggable
    checkpoint engine: create/save/load/commit) while the DeepSpeed engine
    saves or loads.

    Attributes:
        state_dict: state name -> captured dict.
        paths: state name -> path DeepSpeed wanted to write.
    """

    def __init__(self, storage: CheckpointStorage):
        self.state_dict: Dict[str, object] = {}
        self.paths: Dict[str, str] = {}
        self.safetensors_metadata: Dict[str, object] = {}
        self.storage = storage

    def create(self, tag):
        pass

    def save(self, state_dict, path, **kwargs):
        if not isinstance(path, str):  # file objects etc.: not ours
            torch_native_save(state_dict, path)
            return
        name = _state_name(path)
        if name:
            self.state_dict[name] = state_dict
            self.paths[name] = path

    def load(self, path, map_location=None, **kwargs):
        name = _state_name(path, strict=True) if isinstance(path, str) else ""
        if name and name in self.state_dict:
            return self.state_dict[name]
        return self.storage.read_state_dict(
            path, lambda p: torch_native_load(p, map_location=map_location))

    def commit(self, tag):
        pass


class DeepSpeedCheckpointer(Checkpointer):
    """Saves / loads a DeepSpeedEngine through flash checkpoint.

    Example::
        engine, *_ = deepspeed.initialize(...)
        ckpt = DeepSpeedCheckpointer(engine, save_dir)
        ckpt.save_checkpoint(save_dir, tag, storage_type=StorageType.MEMORY)
        ckpt.save_checkpoint(save_dir, tag, storage_type=StorageType.DISK)
        ckpt.load_checkpoint(save_dir)
    """

    def __init__(self, engine, checkpoint_dir, comm_backend="", deletion_strategy=None,
                 save_timeout=CheckpointConstant.SAVE_TIMEOUT, async_drain=None):
        self.engine = engine
        self.checkpoint_dir = checkpoint_dir
        global_shard_num = 1
        if engine.zero_optimization():
            global_shard_num = dist.get_world_size(engine.optimizer.dp_process_group)
        zero_stage = engine.zero_optimization_stage()
        self.storage = get_checkpoint_storage(deletion_strategy)
        self._async_save_engine = DeepSpeedCheckpointEngine(
            checkpoint_dir, storage=self.storage, global_shard_num=global_shard_num,
            zero_stage=zero_stage, comm_backend=comm_backend, save_timeout=save_timeout,
            async_drain=async_drain)
        self._ckpt_agent = AsyncCheckpointAgent(self._async_save_engine.storage)
        self._local_rank = env_utils.get_local_rank()
        self._ds_tracer_file = os.path.join(checkpoint_dir, DeepSpeedCheckpointSaver.TRACER_FILE)
        self._dlrover_tracer_file = os.path.join(checkpoint_dir,
                                                 CheckpointConstant.TRACER_FILE_NAME)
        if zero_stage < _ZERO_STAGE_WEIGHTS and self._local_rank == 0:
            # below ZeRO-3 the module is replicated: one saver per node
            engine.save_non_zero_checkpoint = True

    def _capture(self, save_dir, tag, client_state, save_latest):
        self._ckpt_agent.state_dict = {}
        self._ckpt_agent.paths = {}
        with patched_torch_save(self._ckpt_agent.save):
            self.engine.save_checkpoint(save_dir, tag, client_state, save_latest)
        return self._ckpt_agent.state_dict, self._ckpt_agent.paths

    def save_checkpoint(self, save_dir, tag=None, client_state={}, save_latest=True,
                        storage_type=StorageType.DISK):
        self._async_save_engine.guard_if_in_place(getattr(self.engine, "optimizer", None))
        if storage_type == StorageType.MEMORY:
            sd, paths = self._capture(save_dir, tag, client_state, save_latest)
            self._async_save_engine.save_to_memory(tag, sd, paths)
            self._update_tracer_file(tag)
        elif storage_type == StorageType.DISK:
            sd, paths = self._capture(save_dir, tag, client_state, save_latest)
            self._async_save_engine.save_to_storage(tag, sd, paths)
        else:
            raise ValueError(f"No support storage type {storage_type}")

    def _update_tracer_file(self, tag):
        """DeepSpeedEngine.save_checkpoint made the tag directory and moved its
        `latest` file although nothing reached storage: undo both."""
        if self.engine.global_rank != 0:
            return
        self.storage.safe_rmtree(os.path.join(self.checkpoint_dir, str(tag)))
        committed = self.storage.read(self._dlrover_tracer_file)
        if committed:
            self.storage.write(committed, self._ds_tracer_file)
        else:
            self.storage.safe_remove(self._ds_tracer_file)

    def load_checkpoint(self, load_dir, tag=None, load_module_strict=True,
                        load_optimizer_states=True, load_lr_scheduler_states=True,
                        load_module_only=False, custom_load_fn=None):
        """Same arguments and return value as DeepSpeedEngine.load_checkpoint;
        state found in shared memory wins over the files."""
        self._ckpt_agent.state_dict = self._async_save_engine.load()
        with patched_torch_load(self._ckpt_agent.load):
            return self.engine.load_checkpoint(
                load_dir=load_dir, tag=tag, load_module_strict=load_module_strict,
                load_optimizer_states=load_optimizer_states,
                load_lr_scheduler_states=load_lr_scheduler_states,
                load_module_only=load_module_only, custom_load_fn=custom_load_fn)

    def wait_latest_checkpoint(self, timeout=1800):
        self._async_save_engine.wait_latest_checkpoint(timeout)

    def wait_memory_save(self, timeout=None):
        return self._async_save_engine.wait_memory_save(timeout)
