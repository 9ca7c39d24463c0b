"""ORACLE (test infrastructure / baseline, not product code).

Restatement of the reference's memory-save hot path as it runs on a GPU box:
  dlrover/python/elastic_agent/torch/ckpt_saver.py:303-333  save_state_dict
  dlrover/python/elastic_agent/torch/ckpt_saver.py:198-231  _traverse_copy_to_shm /
      _write_shared_memory: per leaf tensor
      torch.frombuffer(shm, dtype, count, offset).reshape(shape).copy_(tensor)
i.e. one BLOCKING device->pageable-host cudaMemcpy per tensor, issued from the
training thread, into an mmap'd POSIX shm segment.  bench.py times this as the
`--impl reference` arm and as `cpu_baseline` (kind "port") on the same box and
state_dict as the CUDA path.  The two pickles of the meta tree stand in for the
two SharedDict.set() calls (multi_process.py:635-649) the reference makes per
save; socket round trips (sub-millisecond) are not reproduced.

Parity status: PINNED — tests/test_oracle.py::test_ref_port_matches_golden
checks the segment image this produces against the reference-generated goldens.
"""

from __future__ import annotations

import mmap
import os
import pickle
from collections.abc import Mapping

import _posixshmem
import torch

from . import shm_layout


class RefPortSegment:
    """Pageable POSIX shm segment, as multi_process.py:696-734 creates it."""

    def __init__(self, name: str, size: int):
        self.name = "/" + name.lstrip("/")
        try:
            _posixshmem.shm_unlink(self.name)
        except FileNotFoundError:
            pass
        self.fd = _posixshmem.shm_open(self.name, os.O_CREAT | os.O_EXCL | os.O_RDWR, mode=0o600)
        os.ftruncate(self.fd, size)
        self.mmap = mmap.mmap(self.fd, size)
        self.buf = memoryview(self.mmap)
        self.size = size

    def close(self):
        try:
            self.buf.release()
            self.mmap.close()
        except BufferError:
            pass
        os.close(self.fd)
        try:
            _posixshmem.shm_unlink(self.name)
        except FileNotFoundError:
            pass


def _copy_tree(value, meta, buf):
    """ckpt_saver.py:198-218."""
    it = value.items() if isinstance(value, Mapping) else enumerate(value)
    for k, v in it:
        if isinstance(v, (Mapping, list)):
            _copy_tree(v, meta[k], buf)
        elif torch.is_tensor(v):
            m = meta[k]
            if v.numel() == 0:
                continue
            with torch.no_grad():
                dst = torch.frombuffer(buf, dtype=v.dtype, count=v.numel(),
                                       offset=m.offset).reshape(v.shape)
                dst.copy_(v)  # the blocking per-tensor D2H copy
        else:
            meta[k] = v


class RefPortSaver:
    def __init__(self, name: str):
        self.name = name
        self.segment = None
        self.meta = None

    def save(self, state_dict):
        if self.segment is None:
            self.meta, total = shm_layout.plan_layout(state_dict)
            self.segment = RefPortSegment(self.name, total)
        wire = pickle.dumps(self.meta)  # metadata.set(writing_shm=True)
        _copy_tree(state_dict, self.meta, self.segment.buf)
        wire = pickle.dumps(self.meta)  # metadata.set(writing_shm=False)
        return len(wire)

    def close(self):
        if self.segment is not None:
            self.segment.close()
            self.segment = None
