"""Same module name as the reference's fc/full_ckpt_engine.py
(`FullCheckpointEngine`, fc/full_ckpt_engine.py:33-190); the class lives in engine.py."""

from .engine import DdpCheckpointEngine, FullCheckpointEngine  # noqa: F401

__all__ = ["FullCheckpointEngine", "DdpCheckpointEngine"]
