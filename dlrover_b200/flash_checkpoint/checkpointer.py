"""Same module name as the reference's fc/checkpointer.py (`Checkpointer`,
`StorageType`, fc/checkpointer.py:18-65); the classes live in api.py."""

from .api import Checkpointer, StorageType  # noqa: F401

__all__ = ["Checkpointer", "StorageType"]
