"""Scoped interception of ``torch.save`` / ``torch.load``.

DeepSpeed and Megatron-LM build their checkpoint dicts internally and hand them
to ``torch.save(obj, path)``.  The reference captures them by assigning
``torch.save = recorder`` around the framework's own save call
(dlrover/trainer/torch/flash_checkpoint/deepspeed.py:186-226,
megatron.py:172-208).  These context managers do the same swap but always
restore the original, also when the framework raises.
"""

from __future__ import annotations

import contextlib

import torch

torch_native_save = torch.save
torch_native_load = torch.load


@contextlib.contextmanager
def patched_torch_save(recorder):
    previous = torch.save
    torch.save = recorder
    try:
        yield
    finally:
        torch.save = previous


@contextlib.contextmanager
def patched_torch_load(loader):
    previous = torch.load
    torch.load = loader
    try:
        yield
    finally:
        torch.load = previous
