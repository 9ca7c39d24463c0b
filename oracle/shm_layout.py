"""ORACLE (test infrastructure, not product code).

CPU restatement, in numpy + plain Python, of how the reference serialises a
state_dict into the checkpoint shared-memory segment and reads it back.  Each
function cites the reference lines (dlrover @ 468d632, paths relative to
/root/reference) it restates.  Parity status: PINNED — tests/test_oracle.py
checks this module against tests/golden/*.json|*.bin, which were produced by
running the reference itself (tests/golden/make_golden.py) and against the
reference tests' known-answer values (9640-byte SimpleNet image, 1632-byte
ToyModel image, TensorMeta(numel=100, element_size=4, offset=0)).

torch is used only to look at tensors (shape/dtype/bytes); all arithmetic on
offsets is Python int and all byte movement is numpy.
"""

from __future__ import annotations

from collections.abc import Mapping
from dataclasses import dataclass
from typing import Any, Callable, List, Tuple

import numpy as np
import torch


@dataclass
class OracleTensorMeta:
    """dlrover/python/elastic_agent/torch/ckpt_saver.py:88-94 (TensorMeta)."""

    shape: Tuple[int, ...] = ()
    dtype: Any = None
    element_size: int = 0
    numel: int = 0
    offset: int = 0


def traverse(value, visitor: Callable):
    """ckpt_saver.py:118-133 (_traverse_state_dict): depth-first map over
    Mapping / list (tuples are LEAVES), dict order preserved, new containers."""
    if isinstance(value, Mapping):
        return {k: traverse(v, visitor) for k, v in value.items()}
    if isinstance(value, list):
        return [traverse(v, visitor) for v in value]
    return visitor(value)


def plan_layout(state_dict) -> Tuple[Any, int]:
    """ckpt_saver.py:286-301 (_create_tensor_meta) driven by :308-310.
    offset = running sum of numel*element_size in traversal order, no padding;
    non-tensor leaves pass through unchanged.  Returns (meta_tree, total_bytes)."""
    total = 0

    def visit(v):
        nonlocal total
        if not torch.is_tensor(v):
            return v
        m = OracleTensorMeta(
            shape=tuple(v.shape),
            dtype=v.dtype,
            element_size=v.element_size(),
            numel=v.numel(),
            offset=total,
        )
        total += v.numel() * v.element_size()
        return m

    return traverse(state_dict, visit), total


def tensor_bytes(t: torch.Tensor) -> np.ndarray:
    """The bytes `shm_tensor.copy_(t)` deposits (ckpt_saver.py:227-231): the
    tensor's elements in logical (row-major) order, whatever its strides or
    device."""
    c = t.detach().cpu().contiguous()
    if c.numel() == 0:
        return np.zeros(0, dtype=np.uint8)
    return c.reshape(-1).view(torch.uint8).numpy()


def write_image(state_dict, meta_tree, buf: np.ndarray) -> None:
    """ckpt_saver.py:198-231 (_traverse_copy_to_shm + _write_shared_memory).
    Tensors are copied to buf[offset:offset+nbytes]; zero-numel tensors write
    nothing; NON-tensor leaves overwrite the meta entry in place (:207-208,
    :217-218)."""
    if isinstance(state_dict, Mapping):
        it = state_dict.items()
    elif isinstance(state_dict, list):
        it = enumerate(state_dict)
    else:
        return
    for k, v in it:
        if isinstance(v, (Mapping, list)):
            write_image(v, meta_tree[k], buf)
        elif torch.is_tensor(v):
            m = meta_tree[k]
            if v.numel() == 0:
                continue
            b = tensor_bytes(v)
            buf[m.offset : m.offset + b.size] = b
        else:
            meta_tree[k] = v


def serialize(state_dict) -> Tuple[Any, np.ndarray]:
    """plan + write into a fresh image (what the segment holds after
    SharedMemoryHandler.save_state_dict, ckpt_saver.py:303-333)."""
    meta, total = plan_layout(state_dict)
    buf = np.zeros(total, dtype=np.uint8)
    write_image(state_dict, meta, buf)
    return meta, buf


def read_image(meta_tree, buf: np.ndarray):
    """ckpt_saver.py:136-161 (_read_state_dict_from_shm/_read_tensor_from_buf):
    TensorMeta -> tensor viewing buf at offset (numel==0 -> empty tensor of the
    dtype), anything else passes through."""

    def visit(m):
        if isinstance(m, OracleTensorMeta):
            if m.numel == 0:
                return torch.tensor([], dtype=m.dtype)
            nbytes = m.numel * m.element_size
            raw = torch.from_numpy(buf[m.offset : m.offset + nbytes].copy())
            return raw.view(m.dtype).reshape(m.shape)
        return m

    return traverse(meta_tree, visit)


def flatten_tensor_metas(meta_tree) -> List[OracleTensorMeta]:
    out: List[OracleTensorMeta] = []

    def visit(m):
        if isinstance(m, OracleTensorMeta):
            out.append(m)
        return m

    traverse(meta_tree, visit)
    return out


def pack_ranges(srcs: List[np.ndarray], offsets: List[int], total: int) -> np.ndarray:
    """Flat form used against the C-ABI: range i's bytes land at offsets[i]."""
    buf = np.zeros(total, dtype=np.uint8)
    for b, off in zip(srcs, offsets):
        buf[off : off + b.size] = b
    return buf


# ---- FSDP / DCP item accounting -------------------------------------------------


def dcp_item_offsets(item_sizes: List[int]) -> List[Tuple[int, int]]:
    """dlrover/trainer/torch/flash_checkpoint/fsdp_engine.py:85-107,133-155
    (_write_memory_from_list/_write_item): items are laid out back to back in
    plan order; returns (offset, length) per item."""
    out, off = [], 0
    for n in item_sizes:
        out.append((off, n))
        off += n
    return out
