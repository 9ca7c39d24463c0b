"""Same module name as the reference's fc/megatron_engine.py
(`MegatronCheckpointEngine`, `MegatronDistCheckpointEngine`, fc/megatron_engine.py:28-278); the classes
live in engine.py."""

from .engine import MegatronCheckpointEngine, MegatronDistCheckpointEngine  # noqa: F401

__all__ = ["MegatronCheckpointEngine", "MegatronDistCheckpointEngine"]
