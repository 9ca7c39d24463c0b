"""Generates tests/golden/*.bin|*.json by running the REFERENCE implementation
(/root/reference, dlrover @ 468d632) on the fixtures of fixtures.py.

Run here (the reference is not on the GPU box):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
For each fixture it records
  <name>.bin   the bytes the reference's SharedMemoryHandler.save_state_dict
               (ckpt_saver.py:303-333) left in the shm segment
  <name>.json  the meta tree (TensorMeta -> dict), total size, sha256 of the
               image, and sha256 of the file the reference's
               DdpCheckpointSaver.persist_to_storage (ckpt_saver.py:1079-1122)
               writes from that segment with torch.save (+ the torch version,
               since the zip/pickle bytes are torch-version specific).
"""

import hashlib
import json
import os
import sys
import tempfile
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("kubernetes", mock.MagicMock())
os.environ["ROLE_NAME"] = "dlrover-trainer"
os.environ.setdefault("TORCHELASTIC_RUN_ID", f"golden{os.getpid()}")

import torch  # noqa: E402
from dlrover.python.elastic_agent.torch import ckpt_saver as ref  # noqa: E402

import fixtures  # noqa: E402


def meta_to_json(m):
    if isinstance(m, ref.TensorMeta):
        return {"__tensor__": True, "shape": list(m.shape), "dtype": str(m.dtype),
                "element_size": m.element_size, "numel": m.numel, "offset": m.offset}
    if isinstance(m, ref.CheckpointConfig):
        return {"__config__": True, "step": m.step, "rank": m.rank, "paths": m.paths}
    if isinstance(m, dict):
        return {"__dict__": [[repr(k) if not isinstance(k, str) else k, meta_to_json(v)]
                             for k, v in m.items()]}
    if isinstance(m, list):
        return {"__list__": [meta_to_json(v) for v in m]}
    if isinstance(m, tuple):
        return {"__tuple__": list(m)}
    return {"__leaf__": repr(m)}


def main():
    out = {}
    for i, (name, build) in enumerate(fixtures.FIXTURES.items()):
        handler = ref.SharedMemoryHandler(100 + i, host=True)
        tmp = tempfile.mkdtemp()
        path = os.path.join(tmp, "rank_0.pt")
        sd = {"model_states": build()}
        sd[ref.DLROVER_CKPT_CONFIG_KEY] = ref.CheckpointConfig(
            step=5, paths={"model_states": path})
        handler.save_state_dict(sd)
        image = bytes(handler.shared_memory.buf)
        meta = handler.metadata.get(local=True)
        # what the agent would persist from this segment
        loaded = handler.load_state_dict()
        loaded.pop(ref.DLROVER_CKPT_CONFIG_KEY)
        torch.save(loaded["model_states"], path)
        file_sha = hashlib.sha256(open(path, "rb").read()).hexdigest()
        with open(os.path.join(HERE, f"{name}.bin"), "wb") as f:
            f.write(image)
        info = {"size": len(image), "sha256": hashlib.sha256(image).hexdigest(),
                "meta": meta_to_json(meta), "torch_save_sha256": file_sha,
                "torch_version": torch.__version__}
        with open(os.path.join(HERE, f"{name}.json"), "w") as f:
            json.dump(info, f, indent=1)
        out[name] = info["size"]
        del loaded
        handler.shared_memory.unlink()
    print(out)
    for k, v in fixtures.KNOWN_SIZES.items():
        assert out[k] == v, (k, out[k], v)


if __name__ == "__main__":
    main()
