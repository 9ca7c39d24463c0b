"""Flash checkpoint entry points for Megatron-LM training scripts:
drop-in replacements of megatron.checkpointing.save_checkpoint/load_checkpoint.

Reference @ 468d632: dlrover/trainer/torch/flash_checkpoint/megatron.py —
MegatronCheckpointer (:54-136) records Megatron's torch.save calls
("…model_optim_rng.pt" -> "model_states", "…distrib_optim.pt" ->
"optim_states"), save_checkpoint (:139-215), load_checkpoint (:218-247),
wait_latest_checkpoint (:250-260).

Megatron-LM is resolved lazily; tests (and users with forks) may rebind the
module attributes `megatron_save`, `megatron_load`, `get_args`.
"""

from __future__ import annotations

import inspect
import os

import torch.distributed as dist

from ..ckpt_saver import MegatronCheckpointSaver
from ..common.constants import CheckpointConstant
from ..common.log import default_logger as logger
from ..common.singleton import Singleton
from ..common.storage import PosixDiskStorage
from .api import StorageType
from .engine import MegatronCheckpointEngine
from .torch_io_patch import (
    patched_torch_load,
    patched_torch_save,
    torch_native_load,
    torch_native_save,
)

_MODEL_SD_NAME = "model_optim_rng.pt"
_DIST_OPTIM_SD_NAME = "distrib_optim.pt"

try:
    from megatron.training import get_args  # type: ignore
    from megatron.training.checkpointing import load_checkpoint as megatron_load  # type: ignore
    from megatron.training.checkpointing import save_checkpoint as megatron_save  # type: ignore
except ImportError:
    try:
        from megatron import get_args  # type: ignore
        from megatron.checkpointing import load_checkpoint as megatron_load  # type: ignore
        from megatron.checkpointing import save_checkpoint as megatron_save  # type: ignore
    except ImportError:
        get_args = megatron_load = megatron_save = None
        logger.debug("Megatron-LM is not importable; bind megatron_save/megatron_load/get_args.")


def _get_rank():
    return dist.get_rank() if dist.is_initialized() else 0


def _state_name(path: str) -> str:
    if path.endswith(_MODEL_SD_NAME):
        return CheckpointConstant.MODEL_STATES_NAME
    if path.endswith(_DIST_OPTIM_SD_NAME):
        return CheckpointConstant.OPTIM_STATES_NAME
    return ""


class MegatronCheckpointer(Singleton):
    def __init__(self, checkpoint_dir, storage=None, comm_backend="",
                 save_timeout=CheckpointConstant.SAVE_TIMEOUT, replica_count=0,
                 async_drain=None):
        self.state_dict = {}
        self.paths = {}
        self.checkpoint_dir = checkpoint_dir
        self.storage = storage if storage else PosixDiskStorage()
        self.engine = MegatronCheckpointEngine(
            checkpoint_dir=checkpoint_dir, storage=self.storage, comm_backend=comm_backend,
            save_timeout=save_timeout, replica_count=replica_count, async_drain=async_drain)

    def save(self, state_dict, path, **kwargs):
        """torch.save stand-in while Megatron saves."""
        if not isinstance(path, str):
            torch_native_save(state_dict, path)
            return
        name = _state_name(path)
        if not name:
            raise ValueError("MegatronCheckpointer only support the path whose suffix is "
                             f"{_MODEL_SD_NAME} or {_DIST_OPTIM_SD_NAME}.")
        self.state_dict[name] = state_dict
        self.paths[name] = path

    def load(self, path, **kwargs):
        """torch.load stand-in while Megatron loads: shared memory first."""
        if not isinstance(path, str):
            return torch_native_load(path)
        loaded = self.engine.load(resume_path=path)
        in_memory = loaded[1] if isinstance(loaded, tuple) else loaded
        name = _state_name(path)
        if name and in_memory and name in in_memory:
            return in_memory[name]
        return self.storage.read_state_dict(
            path, lambda p: torch_native_load(p, map_location="cpu"))

    def update_tracer_file(self, iteration: int):
        """Megatron's save made iter_XXXXXXX/ and bumped its tracker although
        nothing reached storage yet: remove the directory and put the tracker
        back to the last step the agent really committed."""
        self.storage.safe_rmtree(
            os.path.join(self.checkpoint_dir, "iter_{:07d}".format(iteration)))
        megatron_tracker = os.path.join(self.checkpoint_dir, MegatronCheckpointSaver.TRACER_FILE)
        committed = self.storage.read(
            os.path.join(self.checkpoint_dir, CheckpointConstant.TRACER_FILE_NAME))
        if committed:
            self.storage.write(committed, megatron_tracker)
        else:
            self.storage.safe_remove(megatron_tracker)


def _run_megatron_save(iteration, model, optimizer, opt_param_scheduler, flops):
    if "num_floating_point_operations_so_far" in inspect.signature(megatron_save).parameters:
        megatron_save(iteration, model, optimizer, opt_param_scheduler, flops)
    else:
        megatron_save(iteration, model, optimizer, opt_param_scheduler)


def save_checkpoint(iteration, model, optimizer, opt_param_scheduler,
                    num_floating_point_operations_so_far=0, storage_type=StorageType.DISK,
                    storage=None, comm_backend="", save_timeout=CheckpointConstant.SAVE_TIMEOUT,
                    replica_count=0):
    """Same leading arguments as megatron.checkpointing.save_checkpoint.

    storage: CheckpointStorage (default PosixDiskStorage); comm_backend: backend
    of the control group; replica_count: in-memory replicas on other nodes."""
    if storage_type not in (StorageType.MEMORY, StorageType.DISK):
        raise ValueError(f"No support storage type {storage_type}")
    saver = MegatronCheckpointer.singleton_instance(
        get_args().save, storage=storage, comm_backend=comm_backend, save_timeout=save_timeout,
        replica_count=replica_count)
    try:
        with patched_torch_save(saver.save):
            _run_megatron_save(iteration, model, optimizer, opt_param_scheduler,
                               num_floating_point_operations_so_far)
    finally:
        if _get_rank() == 0:
            saver.update_tracer_file(iteration)
    saver.engine.guard_if_in_place(optimizer)
    if storage_type == StorageType.MEMORY:
        saver.engine.save_to_memory(iteration, saver.state_dict, saver.paths)
    else:
        saver.engine.save_to_storage(iteration, saver.state_dict, saver.paths)
    # the engine keeps its own reference until the gather kernel has run
    saver.state_dict = {}


def load_checkpoint(model, optimizer, opt_param_scheduler, load_arg="load", strict=True,
                    storage=None, comm_backend="", save_timeout=CheckpointConstant.SAVE_TIMEOUT,
                    replica_count=0):
    """Same leading arguments as megatron.checkpointing.load_checkpoint; state
    in shared memory is used before the files."""
    checkpointer = MegatronCheckpointer.singleton_instance(
        get_args().save, storage=storage, comm_backend=comm_backend, save_timeout=save_timeout,
        replica_count=replica_count)
    with patched_torch_load(checkpointer.load):
        return megatron_load(model, optimizer, opt_param_scheduler, load_arg, strict)


def wait_latest_checkpoint(timeout=1800):
    checkpointer = MegatronCheckpointer.singleton_instance(checkpoint_dir=get_args().save)
    checkpointer.engine.wait_latest_checkpoint(timeout)
