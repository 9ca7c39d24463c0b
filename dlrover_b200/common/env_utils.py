"""Process-topology environment readers (torchrun / dlrover-run export these).

Reference: dlrover/python/common/env_utils.py:25-81.  Defaults are part of the
behaviour — note GROUP_RANK defaults to 1 there (env_utils.py:55-56), which the
engines record into CheckpointConfig.group_rank.
"""

import os

from .constants import NodeEnv


def _int_env(key: str, default: int) -> int:
    return int(os.getenv(key, default))


def get_node_rank() -> int:
    if NodeEnv.NODE_RANK in os.environ:
        return int(os.environ[NodeEnv.NODE_RANK])
    return _int_env(NodeEnv.WORKER_RANK, 0)


def get_node_num() -> int:
    return _int_env(NodeEnv.NODE_NUM, 0)


def get_local_world_size() -> int:
    return _int_env("LOCAL_WORLD_SIZE", 1)


def get_local_rank() -> int:
    return _int_env("LOCAL_RANK", 0)


def get_rank() -> int:
    return _int_env("RANK", 0)


def get_group_world_size() -> int:
    return _int_env("GROUP_WORLD_SIZE", 1)


def get_group_rank() -> int:
    return _int_env("GROUP_RANK", 1)


def get_torch_restart_count() -> int:
    return _int_env("TORCHELASTIC_RESTART_COUNT", 0)


def get_run_id() -> str:
    return os.getenv(NodeEnv.TORCHELASTIC_RUN_ID, "")


def launched_by_dlrover_run() -> bool:
    """True when the elastic agent (which hosts the saver threads) started this
    process; engines then must not fork their own saver daemon
    (dlrover/trainer/torch/flash_checkpoint/engine.py:128-137)."""
    return os.getenv(NodeEnv.ROLE_NAME, "") == NodeEnv.TRAINER_ROLE
