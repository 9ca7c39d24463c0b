"""ORACLE (test infrastructure / baseline, not product code): the REFERENCE ITSELF.

Loads the reference's own byte code from oracle/_ref/pyref/ (compiled from
/root/reference by oracle/build_ref.py; nothing of it is in the repository) and drives its
memory-save hot path unmodified:

  dlrover/python/elastic_agent/torch/ckpt_saver.py:245-262   SharedMemoryHandler(local_rank, host=True)
  dlrover/python/elastic_agent/torch/ckpt_saver.py:303-333   save_state_dict
  dlrover/python/elastic_agent/torch/ckpt_saver.py:335-366   load_state_dict

bench.py times `RealRefSaver.save` as the `--impl reference` arm and as `cpu_baseline`
(kind "reference"); when pyref/ is absent or does not import, bench.py falls back to the
restatement oracle/ref_port.py (kind "port") and says so.  tests/test_oracle.py checks that
this, ref_port.py and shm_layout.py leave the same bytes in the segment.

`handler_rank` picks the names of the segment and of the SharedDict socket
(ckpt_shm_<rank>, ckpt_meta_<rank> under /tmp/ckpt_sock/<TORCHELASTIC_RUN_ID>/): callers
use ranks no product handler of the same process uses.
"""

from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PYREF = os.path.join(HERE, "_ref", "pyref")
_MODULE = None


def available() -> bool:
    return os.path.exists(os.path.join(PYREF, "dlrover", "python", "elastic_agent", "torch",
                                       "ckpt_saver.pyc"))


def load():
    """The reference's ckpt_saver module, imported from pyref/ (never from /root/reference:
    the GPU box has none)."""
    global _MODULE
    if _MODULE is not None:
        return _MODULE
    if not available():
        raise ImportError(f"{PYREF} not built (python oracle/build_ref.py, needs /root/reference)")
    mine = os.path.realpath(PYREF) + os.sep
    for name, mod in list(sys.modules.items()):
        if name == "dlrover" or name.startswith("dlrover."):
            f = os.path.realpath(getattr(mod, "__file__", None) or "")
            if not f.startswith(mine):
                raise ImportError(f"{name} is already imported from {f or '?'} (compat aliases?); "
                                  "the reference needs a process of its own")
    os.environ.setdefault("ROLE_NAME", "dlrover-trainer")
    sys.path.insert(0, PYREF)
    try:
        import logging

        from dlrover.python.elastic_agent.torch import ckpt_saver as ref
        # two INFO lines per save on stderr are not what is being timed
        logging.getLogger("dlrover").setLevel(logging.WARNING)
        try:
            from dlrover.python.common.log import default_logger
            default_logger.setLevel(logging.WARNING)
        except Exception:
            pass
    except Exception:
        sys.path.remove(PYREF)
        for name in [n for n in sys.modules if n == "dlrover" or n.startswith("dlrover.")]:
            del sys.modules[name]
        raise
    f = os.path.realpath(ref.__file__)
    if not f.startswith(mine):
        raise ImportError(f"dlrover resolved to {f}, not to {PYREF}")
    _MODULE = ref
    return ref


class RealRefSaver:
    """Same surface as oracle/ref_port.RefPortSaver, the reference's handler underneath."""

    kind = "reference"

    def __init__(self, handler_rank: int):
        self.ref = load()
        self.handler = self.ref.SharedMemoryHandler(handler_rank, host=True)
        self.step = 0

    def save(self, state_dict):
        """`state_dict` without the config entry; it is added the way the reference's
        engine does (engine.py:364-376)."""
        self.step += 1
        sd = dict(state_dict)
        sd[self.ref.DLROVER_CKPT_CONFIG_KEY] = self.ref.CheckpointConfig(step=self.step, paths={})
        self.handler.save_state_dict(sd)

    @property
    def buf(self):
        return self.handler.shared_memory.buf

    def views(self):
        """The reference's restore: tensors aliasing the segment (ckpt_saver.py:144-161)."""
        return self.handler.load_state_dict()

    def close(self):
        try:
            self.handler.unlink()
        finally:
            try:
                self.handler.close()
            except BufferError:   # a view of the segment is still alive somewhere
                pass
