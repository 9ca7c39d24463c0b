"""Names that are part of the on-disk / inter-process contract.

Reference: dlrover/python/common/constants.py:444-448 (CheckpointConstant),
:326-356 (NodeEnv, only the keys this path reads), :509-542 (event actions).
The values are format, not code: tracker file names, state names and env keys
must match for a checkpoint written by one implementation to be readable by
the other.
"""


class CheckpointConstant:
    TRACER_FILE_NAME = "dlrover_latest.txt"
    MODEL_STATES_NAME = "model_states"
    OPTIM_STATES_NAME = "optim_states"
    SAVE_TIMEOUT = 600


class NodeEnv:
    NODE_NUM = "NODE_NUM"
    NODE_RANK = "NODE_RANK"
    WORKER_RANK = "WORKER_RANK"  # pre-0.3.0 spelling of NODE_RANK
    TORCHELASTIC_RUN_ID = "TORCHELASTIC_RUN_ID"
    ROLE_NAME = "ROLE_NAME"
    TRAINER_ROLE = "dlrover-trainer"


class EventReportConstants:
    TYPE_INFO = "info"
    TYPE_WARN = "warn"
    TYPE_ERROR = "error"
    ACTION_SAVE_SHARD_START = "save_shard_start"
    ACTION_SAVE_SHARD_COMPLETE = "save_shard_complete"
    ACTION_SAVE_SHARD_ERROR = "save_shard_error"
    ACTION_MEM_CKPT_START = "mem_ckpt_start"
    ACTION_MEM_CKPT_COMPLETE = "mem_ckpt_complete"
    ACTION_RESUME_MEM_CKPT_START = "resume_mem_ckpt_start"
    ACTION_RESUME_MEM_CKPT_COMPLETE = "resume_mem_ckpt_complete"


class TrainingExceptionLevel:
    PROCESS_ERROR = "process_error"
