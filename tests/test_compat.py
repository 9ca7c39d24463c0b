"""Reference import paths resolve to this package after compat.install()."""

import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import sys
sys.path.insert(0, %r)
import dlrover_b200.compat as compat
assert compat.install() is True
from dlrover.trainer.torch.flash_checkpoint.ddp import DdpCheckpointer
from dlrover.trainer.torch.flash_checkpoint.checkpointer import Checkpointer, StorageType
from dlrover.trainer.torch.flash_checkpoint.full_ckpt_engine import FullCheckpointEngine
from dlrover.trainer.torch.flash_checkpoint.engine import CheckpointEngine, check_all_rank_ready
from dlrover.trainer.torch.flash_checkpoint.deepspeed_engine import DeepSpeedCheckpointEngine
from dlrover.trainer.torch.flash_checkpoint.megatron_engine import MegatronCheckpointEngine
from dlrover.trainer.torch.flash_checkpoint.fsdp_engine import FsdpCheckpointEngine, SharedMemoryWriter
from dlrover.trainer.torch.flash_checkpoint.fsdp import FsdpShardCheckpointer, FsdpFullCheckpointer
from dlrover.trainer.torch.flash_checkpoint import megatron, megatron_dist_ckpt
from dlrover.trainer.torch.flash_checkpoint.deepspeed import DeepSpeedCheckpointer
from dlrover.python.elastic_agent.torch.ckpt_saver import (AsyncCheckpointSaver, SharedMemoryHandler,
    TensorMeta, CheckpointConfig, DdpCheckpointSaver, FsdpDcpSaver, DLROVER_CKPT_CONFIG_KEY)
from dlrover.python.common.multi_process import SharedLock, SharedQueue, SharedDict, SharedMemory
from dlrover.python.common.storage import PosixDiskStorage, KeepLatestStepStrategy
from dlrover.python.common.constants import CheckpointConstant
import dlrover_b200.flash_checkpoint.api as api
assert DdpCheckpointer is api.DdpCheckpointer and issubclass(DdpCheckpointer, Checkpointer)
assert DLROVER_CKPT_CONFIG_KEY == "_DLORVER_CKPT_CONFIG"
assert CheckpointConstant.TRACER_FILE_NAME == "dlrover_latest.txt"
assert hasattr(megatron, "save_checkpoint") and hasattr(megatron_dist_ckpt, "get_parameter_state")
megatron.get_args = lambda: 1          # monkeypatching through the alias reaches the module
import dlrover_b200.flash_checkpoint.megatron as real
assert real.get_args() == 1
print("COMPAT_OK")
"""


def test_aliases_in_clean_interpreter():
    env = dict(os.environ, DLROVER_LOG_LEVEL="ERROR")
    env.pop("PYTHONPATH", None)
    out = subprocess.run([sys.executable, "-c", CODE % ROOT], capture_output=True, text=True,
                         env=env, timeout=300)
    assert "COMPAT_OK" in out.stdout, out.stderr[-2000:]
