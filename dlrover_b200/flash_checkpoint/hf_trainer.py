"""Flash checkpoint for HuggingFace `transformers.Trainer`.

Reference @ 468d632: dlrover/trainer/torch/flash_checkpoint/hf_trainer.py —
HfFlashCheckpointer / HfDdpCheckpointer / HfDeepSpeedCheckpointer (:59-117) and
FlashCkptTrainer (:119-388): `Trainer._save_checkpoint` is re-implemented there
with torch.save swapped for a recorder, against a 2023 transformers.

Here the same effect is obtained without re-stating Trainer internals (they
have drifted: transformers 5.x has no `save_function`/`safe_serialization`
arguments and always writes safetensors): the STOCK `_save_checkpoint` runs
while `torch.save` and transformers' `safe_save_file` are redirected to the
recorder, so whatever tensors the installed Trainer would have written
(optimizer.pt, scheduler.pt, rng_state*.pth, model*.safetensors /
pytorch_model*.bin) land in the flash-checkpoint engine instead, keyed by their
file names, with their intended paths; json/config files are written by the
Trainer as usual.  The agent persists each entry to its path — with
`safetensors.save_file` for entries recorded from `safe_save_file` (the
"safe_serialization" marker the reference's saver already understands,
ckpt_saver.py:1088-1120) and `torch.save` otherwise.

`FlashCkptTrainer` is created on first attribute access so importing this
module does not import transformers.
"""

from __future__ import annotations

import os
import re
import shutil
from typing import Optional

import torch.distributed as dist

from ..common.log import default_logger as logger
from ..common.storage import PosixDiskStorage
from .deepspeed import AsyncCheckpointAgent
from .engine import CheckpointEngine, DeepSpeedCheckpointEngine, FullCheckpointEngine
from .torch_io_patch import patched_torch_save

PREFIX_CHECKPOINT_DIR = "checkpoint"


class HfFlashCheckpointer:
    def __init__(self, checkpoint_dir, storage=None):
        self.checkpoint_dir = checkpoint_dir
        self.storage = storage if storage else PosixDiskStorage()
        self.ckpt_agent = AsyncCheckpointAgent(self.storage)
        self.async_save_engine: Optional[CheckpointEngine] = None

    def save_checkpoint_to_memory(self, step, blocking=False):
        return self.async_save_engine.save_to_memory(step, self.ckpt_agent.state_dict,
                                                     self.ckpt_agent.paths, blocking)

    def save_checkpoint_to_storage(self, step, blocking=False):
        return self.async_save_engine.save_to_storage(step, self.ckpt_agent.state_dict,
                                                      self.ckpt_agent.paths, blocking)


class HfDeepSpeedCheckpointer(HfFlashCheckpointer):
    def __init__(self, engine, checkpoint_dir, storage=None, comm_backend=""):
        super().__init__(checkpoint_dir, storage)
        self.engine = engine
        global_shard_num = 1
        if engine.zero_optimization():
            global_shard_num = dist.get_world_size(engine.optimizer.dp_process_group)
        self.async_save_engine = DeepSpeedCheckpointEngine(
            checkpoint_dir, storage=self.storage, global_shard_num=global_shard_num,
            zero_stage=engine.zero_optimization_stage(), comm_backend=comm_backend)


class HfDdpCheckpointer(HfFlashCheckpointer):
    def __init__(self, checkpoint_dir, storage=None, comm_backend=""):
        super().__init__(checkpoint_dir, storage)
        self.async_save_engine = FullCheckpointEngine(checkpoint_dir, storage=self.storage,
                                                      comm_backend=comm_backend)


class _SafetensorsRecorder:
    """Stands in for transformers.modeling_utils.safe_save_file."""

    def __init__(self, agent: AsyncCheckpointAgent):
        self.agent = agent

    def __call__(self, tensors, filename, metadata=None):
        name = os.path.basename(str(filename))
        self.agent.state_dict[name] = dict(tensors)
        self.agent.paths[name] = str(filename)
        # marker understood by the agent-side saver (safetensors for these)
        self.agent.state_dict["safe_serialization"] = True
        self.agent.safetensors_metadata[name] = metadata


def _build_trainer_class():
    import transformers
    from transformers import Trainer

    class FlashCkptTrainer(Trainer):
        """Trainer whose checkpoints go to shared memory synchronously (cheap)
        and to storage asynchronously through the flash-checkpoint agent.

        The checkpoint directory is created at once with the json/config files;
        weights and optimizer state arrive later.  The last COMPLETE checkpoint
        is the step in `<output_dir>/dlrover_latest.txt`
        (`get_last_checkpoint()`)."""

        def _flash_checkpointer(self, run_dir):
            if not hasattr(self, "flash_checkpointer"):
                if self.is_deepspeed_enabled:
                    self.flash_checkpointer = HfDeepSpeedCheckpointer(self.model_wrapped, run_dir)
                elif not self.is_fsdp_enabled:
                    self.flash_checkpointer = HfDdpCheckpointer(run_dir)
                else:
                    raise ValueError("Flash Checkpoint only supports DeepSpeed or DDP.")
            return self.flash_checkpointer

        def _save_checkpoint(self, model, trial, *args, **kwargs):
            import transformers.modeling_utils as mu

            run_dir = self._get_output_dir(trial=trial)
            step = self.state.global_step
            output_dir = os.path.join(run_dir, f"{PREFIX_CHECKPOINT_DIR}-{step}")
            ckpt = self._flash_checkpointer(run_dir)
            agent = ckpt.ckpt_agent
            agent.state_dict, agent.paths = {}, {}
            agent.safetensors_metadata = {}
            recorder = _SafetensorsRecorder(agent)
            native_safe = mu.safe_save_file
            mu.safe_save_file = recorder
            # rotation must not count directories whose tensors are not on
            # storage yet: done below with our own rule
            limit, self.args.save_total_limit = self.args.save_total_limit, None
            try:
                with patched_torch_save(agent.save):
                    super()._save_checkpoint(model, trial, *args, **kwargs)
            finally:
                mu.safe_save_file = native_safe
                self.args.save_total_limit = limit
            success = ckpt.save_checkpoint_to_storage(step)
            if not success:
                logger.info(f"Skip saving the checkpoint of step {step} because the latest "
                            "checkpoint is not finished.")
                shutil.rmtree(output_dir, ignore_errors=True)
            if self.args.should_save:
                self._rotate_flash_checkpoints(run_dir)

        def _rotate_flash_checkpoints(self, run_dir):
            """Delete the oldest COMPLETE checkpoints beyond save_total_limit
            (directories newer than the tracker step may still be filling)."""
            limit = self.args.save_total_limit
            if limit is None or limit <= 0:
                return
            last = self._get_last_checkpoint_step()
            done = []
            for name in os.listdir(run_dir):
                m = re.fullmatch(rf"{PREFIX_CHECKPOINT_DIR}-([0-9]+)", name)
                if m and int(m.group(1)) <= last:
                    done.append((int(m.group(1)), os.path.join(run_dir, name)))
            done.sort()
            if len(done) <= limit:
                return
            best = self.state.best_model_checkpoint
            if best is not None and limit == 1 and done[-1][1] != best:
                limit = 2
            for _, path in done[:max(0, len(done) - limit)]:
                if path == best:
                    continue
                logger.info(f"Deleting older checkpoint [{path}] due to save_total_limit = "
                            f"{self.args.save_total_limit}.")
                shutil.rmtree(path, ignore_errors=True)

        def get_last_checkpoint(self):
            step = self._get_last_checkpoint_step()
            if step == 0:
                return False
            return os.path.join(self.args.output_dir, f"{PREFIX_CHECKPOINT_DIR}-{step}")

        def _get_last_checkpoint_step(self):
            tracker = os.path.join(self.args.output_dir, "dlrover_latest.txt")
            if not os.path.exists(tracker):
                return 0
            with open(tracker, "r") as f:
                return int(f.read())

        def wait_latest_checkpoint(self, timeout=1800):
            self.flash_checkpointer.async_save_engine.wait_latest_checkpoint(timeout)

    FlashCkptTrainer.__transformers_version__ = transformers.__version__
    return FlashCkptTrainer


_trainer_cls = None


def __getattr__(name):
    global _trainer_cls
    if name == "FlashCkptTrainer":
        if _trainer_cls is None:
            _trainer_cls = _build_trainer_class()
        return _trainer_cls
    raise AttributeError(name)
