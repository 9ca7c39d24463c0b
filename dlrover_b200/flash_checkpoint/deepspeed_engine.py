"""Same module name as the reference's fc/deepspeed_engine.py
(`DeepSpeedCheckpointEngine`, fc/deepspeed_engine.py:31-162); the class lives in engine.py."""

from .engine import DeepSpeedCheckpointEngine  # noqa: F401

__all__ = ["DeepSpeedCheckpointEngine"]
