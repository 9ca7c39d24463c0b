"""Golden for the DCP-over-shm path: runs the REFERENCE's FsdpCheckpointEngine
(dlrover/trainer/torch/flash_checkpoint/fsdp_engine.py @ 468d632, world_size 1,
gloo) on the state dict of fixtures.fixture_dcp() and records the segment image
and the per-item (fqn, offset, length) table of its DCP metadata.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_fsdp.py
"""

import hashlib
import json
import os
import sys
import tempfile
import time
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("kubernetes", mock.MagicMock())
os.environ["ROLE_NAME"] = "dlrover-trainer"
os.environ["TORCHELASTIC_RUN_ID"] = f"goldenfsdp{os.getpid()}"
os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29871", "RANK": "0",
                   "WORLD_SIZE": "1", "LOCAL_RANK": "0", "LOCAL_WORLD_SIZE": "1"})

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from dlrover.python.common.storage import PosixDiskStorage  # noqa: E402
from dlrover.python.elastic_agent.torch.ckpt_saver import AsyncCheckpointSaver  # noqa: E402
from dlrover.trainer.torch.flash_checkpoint.fsdp_engine import FsdpCheckpointEngine  # noqa: E402

import fixtures  # noqa: E402

dist.init_process_group("gloo", rank=0, world_size=1)
AsyncCheckpointSaver.start_async_saving_ckpt()
tmp = tempfile.mkdtemp()
engine = FsdpCheckpointEngine(tmp, PosixDiskStorage())
sd = fixtures.fixture_dcp()
assert engine.save_to_memory(7, sd, {"model_states": os.path.join(tmp, "7")})
time.sleep(0.5)
image = bytes(engine._shm_handler.shared_memory.buf)
meta = engine._shm_handler.metadata.get()
dcp = meta["dcp_metadata"]
items = sorted((idx.fqn, list(idx.offset) if idx.offset is not None else None, info.relative_path,
                info.offset, info.length) for idx, info in dcp.storage_data.items())
out = {"size": len(image), "sha256": hashlib.sha256(image).hexdigest(), "items": items,
       "no_shard_keys": sorted(meta["no_shard_data"].keys()),
       "path": os.path.relpath(meta["_DLORVER_CKPT_CONFIG"].paths["model_states"], tmp),
       "torch_version": torch.__version__}
with open(os.path.join(HERE, "dcp_plain.bin"), "wb") as f:
    f.write(image)
with open(os.path.join(HERE, "dcp_plain.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out["size"], out["items"][:3], out["path"])
engine._shm_handler.shared_memory.unlink()
os._exit(0)
