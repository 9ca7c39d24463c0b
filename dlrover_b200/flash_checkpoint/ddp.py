"""Same module name as the reference's fc/ddp.py (`DdpCheckpointer`, fc/ddp.py:25-125);
the class lives in api.py."""

from .api import Checkpointer, DdpCheckpointer, StorageType  # noqa: F401

__all__ = ["Checkpointer", "DdpCheckpointer", "StorageType"]
