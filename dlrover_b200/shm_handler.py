"""Shared-memory handler: serialises a (nested) state_dict into the node-local
POSIX shm segment the agent persists from, and reads it back.

Contract (reference @ 468d632, dlrover/python/elastic_agent/torch/ckpt_saver.py):
  * TensorMeta / CheckpointConfig dataclasses (:88-115) and the config key
    "_DLORVER_CKPT_CONFIG" (:50; the typo is format);
  * layout = traversal order of Mapping/list containers, offset = running sum
    of numel*element_size, NO padding (:118-133, :286-301) — SimpleNet + empty
    SGD state is exactly 9640 bytes;
  * segment name "[<run_id>_]ckpt_shm_<local_shard>" (:247-259), meta tree kept
    in the agent's SharedDict "ckpt_meta_<local_shard>" (:261);
  * save protocol: writing_shm=True -> meta set -> bytes -> writing_shm=False
    -> meta set (:303-333); a reader that sees writing_shm gets {} (:343-347);
  * load returns CPU tensors aliasing the segment (:136-161).

What is different underneath: device-resident leaves are NOT copied one by one
with blocking pageable cudaMemcpy (:221-231).  They are described once to
libflashckpt (a cached plan), gathered by one sm_100a kernel into an HBM arena
on the caller's stream, and drained to the pinned segment by DMA on a side
stream; `save_state_dict(..., blocking=False)` returns as soon as the kernel is
enqueued and finishes the protocol from a completion thread.  Host-resident
leaves go straight into the segment with fc_host_pack.  There is no torch
fallback for CUDA leaves: a missing library raises.
"""

from __future__ import annotations

import atexit
import json
import os
import threading
import time
import weakref
from collections.abc import Mapping
from dataclasses import dataclass
from datetime import datetime
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch

from . import _native as native
from .common.constants import EventReportConstants, NodeEnv
from .common.log import default_logger as logger
from .common.multi_process import SharedDict, SharedMemory

DLROVER_CKPT_CONFIG_KEY = "_DLORVER_CKPT_CONFIG"


def report_local_event(event_type="", instance="", action="", msg="", labels=None):
    stamp = datetime.now().strftime("%Y-%m-%d, %H:%M:%S")
    logger.info(f"[{stamp}][{event_type}][{instance}][{action}][{msg}][{json.dumps(labels or {})}]")


class CheckpointSharedObjPrefix:
    SAVE_STEP_QNAME = "ckpt_lock_rank_"
    META_NAME = "ckpt_meta_"
    SHM_NAME = "ckpt_shm_"
    SHM_LOCK_NAME = "shm_lock_"


@dataclass
class TensorMeta:
    shape: Tuple[int, ...] = None  # type: ignore
    dtype: torch.dtype = None  # type: ignore
    element_size: int = 0
    numel: int = 0
    offset: int = 0


@dataclass
class CheckpointConfig:
    """Per-shard header stored under DLROVER_CKPT_CONFIG_KEY in the meta tree.

    step: global iteration; writing_shm: True while the trainer is filling the
    segment (a reader must not trust the bytes); paths: state name -> storage
    path the agent persists that sub-dict to."""

    rank: int = 0
    group_rank: int = 0
    world_size: int = 0
    step: int = 0
    writing_shm: bool = False
    paths: Dict[str, str] = None  # type: ignore


# ------------------------------------------------------------------- traversal --


def _traverse_state_dict(value, visitor: Callable):
    """Depth-first map over Mapping / list containers (tuples are leaves), dict
    order kept, fresh containers returned."""
    if isinstance(value, Mapping):
        return {k: _traverse_state_dict(v, visitor) for k, v in value.items()}
    if isinstance(value, list):
        return [_traverse_state_dict(v, visitor) for v in value]
    return visitor(value)


class _Layout:
    """Result of one planning walk over a state_dict."""

    __slots__ = ("meta", "total", "device_leaves", "host_leaves", "leaf_metas", "reused")

    def __init__(self):
        self.meta: Any = None
        self.total = 0
        self.device_leaves: List[Tuple[torch.Tensor, TensorMeta]] = []
        self.host_leaves: List[Tuple[torch.Tensor, TensorMeta]] = []
        self.leaf_metas: List[TensorMeta] = []  # every tensor leaf, traversal order
        self.reused = 0  # leaves whose TensorMeta was taken over from `prev`


def plan_layout(state_dict, prev: Optional["_Layout"] = None) -> _Layout:
    """Assign every tensor leaf its byte offset (reference layout,
    ckpt_saver.py:286-301) and sort the leaves into device-resident and
    host-resident.  This runs on the training thread at every save, so the walk
    is kept lean: with `prev` (the layout of the previous save) a leaf that
    still has the same shape/dtype at the same offset takes over its TensorMeta
    instead of building a new one."""
    lay = _Layout()
    metas, dev, host = lay.leaf_metas, lay.device_leaves, lay.host_leaves
    prev_metas = prev.leaf_metas if prev is not None else ()
    nprev = len(prev_metas)
    Tensor = torch.Tensor
    total = 0
    reused = 0

    def walk(value):
        nonlocal total, reused
        if isinstance(value, Tensor):
            i = len(metas)
            m = prev_metas[i] if i < nprev else None
            if m is not None and m.offset == total and m.dtype is value.dtype \
                    and m.shape == value.shape:
                reused += 1
            else:
                m = TensorMeta(shape=tuple(value.shape), dtype=value.dtype,
                               element_size=value.element_size(), numel=value.numel(),
                               offset=total)
            metas.append(m)
            nbytes = m.numel * m.element_size
            if nbytes:
                total += nbytes
                (dev if value.is_cuda else host).append((value, m))
            return m
        kind = type(value)
        if kind is dict or (kind is not list and isinstance(value, Mapping)):
            return {k: walk(v) for k, v in value.items()}
        if kind is list or isinstance(value, list):
            return [walk(v) for v in value]
        return value  # non-tensor leaf (tuples included): carried in the meta tree

    lay.meta = walk(state_dict)
    lay.total = total
    lay.reused = reused
    return lay


def _read_tensor_from_buf(value, shm: SharedMemory):
    if not isinstance(value, TensorMeta):
        return value
    if value.numel == 0:
        return torch.tensor([], dtype=value.dtype)
    flat = torch.frombuffer(shm.buf, dtype=value.dtype, offset=value.offset, count=value.numel)
    return flat.reshape(value.shape)


def _read_state_dict_from_shm(meta_dict, shm: SharedMemory):
    return _traverse_state_dict(meta_dict, lambda m: _read_tensor_from_buf(m, shm))


def _create_shared_memory(name, create, size=0):
    """Attach to, or (re)create with the requested size, the named segment."""
    if not create:
        try:
            return SharedMemory(name=name)
        except FileNotFoundError:
            return None
    if size == 0:
        logger.warning("Cannot create the shared memory with size = 0.")
        return None
    try:
        return SharedMemory(name=name, create=True, size=size)
    except FileExistsError:
        old = SharedMemory(name=name)
        if old.size == size:
            return old
        logger.info(f"The old size is {old.size} and create a new memory buffer with size {size}.")
        old.unlink()
        old.close()
        return SharedMemory(name=name, create=True, size=size)


def _host_threads() -> int:
    env = os.getenv("DLROVER_B200_HOST_THREADS", "")
    if env:
        return max(1, int(env))
    # measured on config[0] (GPT-2 small on CPU, tools/cpu_config_bench.py): all cores of an
    # 8-vCPU box give 13.7 ms vs 25 ms with half of them (reference: 16.4 ms)
    return max(1, min(32, os.cpu_count() or 1))


# -------------------------------------------------------------- device staging --


class _DeviceStager:
    """Owns the libflashckpt context usage for ONE segment: host registration,
    arena sizing and the cached plan."""

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.ctx = native.get_context(device_index)
        self._registered_addr = 0
        self._plans: Dict[str, native.Plan] = {}  # role -> plan
        self.register_seconds = 0.0
        self._warned_no_arena = False

    def attach(self, shm: SharedMemory):
        addr = shm.address
        if addr == self._registered_addr:
            return
        self.detach()
        t0 = time.time()
        self.ctx.host_register(addr, shm.size, prefault_threads=_host_threads())
        self.register_seconds = time.time() - t0
        self._registered_addr = addr
        logger.info(
            f"Pinned the {shm.size / 2**30:.2f} GiB checkpoint segment for DMA "
            f"in {self.register_seconds:.2f}s.")

    def detach(self):
        if self._registered_addr:
            try:
                self.ctx.host_unregister(self._registered_addr)
            except native.NativeError as e:
                logger.warning(f"host_unregister: {e}")
            self._registered_addr = 0

    def plan_for(self, ranges: List[Tuple[torch.Tensor, int, int]], keepalive: list,
                 role: str = "save", stream=None):
        """ranges: (tensor, segment offset, nbytes) per device-resident leaf.
        One plan object per role ("save" / "restore"); when the tensors moved
        (FSDP hands out fresh ones on every state_dict()) the plan is
        re-targeted in place with a stream-ordered table upload."""
        if not all(t.is_contiguous() for t, _, _ in ranges):
            # rare (state_dict tensors are contiguous); device-side repack,
            # kept alive until the pack kernel has consumed it
            packed = []
            for t, off, nbytes in ranges:
                if not t.is_contiguous():
                    t = t.detach().contiguous()
                    keepalive.append(t)
                packed.append((t, off, nbytes))
            ranges = packed
        ptrs = [t.data_ptr() for t, _, _ in ranges]
        offs = [r[1] for r in ranges]
        lens = [r[2] for r in ranges]
        key = (ptrs, offs, lens)
        plan = self._plans.get(role)
        if plan is not None and plan.key == key:
            return plan
        if plan is None:
            plan = self._plans[role] = self.ctx.plan(ptrs, offs, lens)
        else:
            plan.update(ptrs, offs, lens, stream)
        return plan

    ARENA_FULL, ARENA_WINDOWED, ARENA_NONE = "full", "windowed", "none"

    def ensure_arena(self, plan) -> str:
        """Make room for the snapshot of `plan` in HBM.
        "full": the arena covers the plan (the training stream then only waits for
        the gather kernel).  "windowed": DLROVER_B200_ARENA_LIMIT_MB caps the arena
        and the checkpoint is streamed through it window by window (blocking).
        "none": HBM has no room for a second copy of the state — the caller saves
        in place instead (DMA straight from the tensors)."""
        end = plan.arena_end
        if end <= self.ctx.arena_info()[1]:
            return self.ARENA_FULL
        forced = int(os.getenv("DLROVER_B200_ARENA_LIMIT_MB", "0") or 0)
        if forced:
            limit = max(8, forced) << 20
            self.ctx.set_arena_limit(limit)
            self.ctx.arena_reserve(min(end, limit))
            return self.ARENA_FULL if end <= limit else self.ARENA_WINDOWED
        try:
            self.ctx.arena_reserve(end)
            return self.ARENA_FULL
        except native.NativeError as e:
            if e.code != native.FC_ENOMEM:
                raise
        if not self._warned_no_arena:
            self._warned_no_arena = True
            free, _ = torch.cuda.mem_get_info(self.device_index)
            logger.warning(
                f"No room for a {end / 2**30:.1f} GiB snapshot arena in HBM ({free / 2**30:.1f} "
                "GiB free): checkpoints are drained in place (DMA straight from the tensors). "
                "Saves block until the data is in shared memory unless the engine runs with "
                "in_place=True and the optimizer is guarded (see README, 'in-place saves').")
        return self.ARENA_NONE

    def close(self):
        for p in self._plans.values():
            p.destroy()
        self._plans.clear()
        self.detach()


class PendingSave:
    """Handle of a save whose drain is still running on the copy stream."""

    def __init__(self, ctx: Optional[native.Context], ticket: int, finish: Callable[[], None],
                 keepalive: list, pre_drain: Optional[Callable[[], None]] = None,
                 on_error: Optional[Callable[[], None]] = None):
        self._on_error = on_error
        self._ctx = ctx
        self._ticket = ticket
        self._finish = finish
        self._keepalive = keepalive
        self._pre_drain = pre_drain  # set <=> the drain is held until we release it
        self._done = threading.Event()
        self._error: Optional[BaseException] = None
        self._lock = threading.Lock()
        self.timings: Optional[Tuple[float, float, float]] = None

    def _complete(self):
        """Wait for the DMA, then close the protocol.  Runs exactly once."""
        with self._lock:
            if self._done.is_set():
                return
            try:
                if self._pre_drain is not None:
                    try:
                        self._pre_drain()  # e.g. publish writing_shm=True to the agent
                    except BaseException:
                        # the agent does not know the segment is about to change: leave
                        # every byte of it alone (the previous checkpoint stays valid)
                        if not self._ctx.save_cancel(self._ticket):
                            self._ctx.save_release(self._ticket)
                            self._ctx.save_wait(self._ticket)
                        raise
                    self._ctx.save_release(self._ticket)
                if self._ctx is not None:
                    self._ctx.save_wait(self._ticket)
                    self.timings = self._ctx.save_timings(self._ticket)
                self._finish()
            except BaseException as e:  # surfaced by wait()
                self._error = e
                logger.error(f"flash checkpoint drain failed: {e}", exc_info=True)
                if self._on_error is not None:
                    try:
                        self._on_error()  # e.g. give the shard lock back
                    except BaseException:
                        logger.error("error handler of the failed drain raised", exc_info=True)
            finally:
                self._keepalive.clear()
                self._done.set()

    def done(self) -> bool:
        return self._done.is_set()

    def wait(self, timeout: Optional[float] = None) -> bool:
        ok = self._done.wait(timeout)
        if ok and self._error is not None:
            raise self._error
        return ok


# --------------------------------------------------------------------- handler --


class SharedMemoryHandler:
    """Writes / reads the state dict of one local shard to / from shared memory.

    Args:
        local_rank: index of the local shard (names the segment and meta dict).
        host: True on the agent (owns the SharedDict server), False in a
            training process.
    """

    def __init__(self, local_rank, host=True):
        self._buffer_size = 0
        self._local_rank = local_rank
        run_id = os.getenv(NodeEnv.TORCHELASTIC_RUN_ID, "")
        base = CheckpointSharedObjPrefix.SHM_NAME + str(local_rank)
        self._shm_name = f"{run_id}_{base}" if run_id else base
        self.shared_memory: Optional[SharedMemory] = None
        self.metadata = SharedDict(name=CheckpointSharedObjPrefix.META_NAME + str(local_rank),
                                   create=host)
        self._need_creation = True
        self._layout: Optional[_Layout] = None  # of the previous save (TensorMeta reuse)
        self._stager: Optional[_DeviceStager] = None
        self._pending: Optional[PendingSave] = None
        self._master_client = None
        self.last_timings: Optional[Tuple[float, float, float]] = None
        # torch.cuda.Event recorded right after the last gather kernel
        self.last_pack_event = None
        # in-place saves (no HBM snapshot): opt-in, see write_ranges
        self.in_place = os.getenv("DLROVER_B200_IN_PLACE", "0") == "1"
        # HBM an in-place save may spend on snapshotting the tail of the state (the rest
        # is drained in place first): shortens the time the tensors stay frozen
        self.in_place_snapshot_bytes = int(os.getenv("DLROVER_B200_IN_PLACE_SNAPSHOT_MB", "0")) << 20
        self.last_save_in_place = False
        self.last_restore_stats: Dict[str, float] = {}
        self.last_hybrid_cut = None
        self._snapshot_ceiling = 0  # largest snapshot arena HBM turned out to have room for
        self._last_ticket = None
        if not host:
            # a drain still in flight when the interpreter exits must finish (the
            # completion thread is a daemon): otherwise the last checkpoint of a
            # run that ends right after save_checkpoint() would stay half written
            ref = weakref.ref(self)
            atexit.register(lambda: ref() is not None and ref()._finish_at_exit())

    # -- lifecycle ------------------------------------------------------------------
    def close(self):
        try:
            self.wait_pending()
        except BaseException as e:  # already logged by the completion thread
            logger.warning(f"closing after a failed drain: {e}")
        if self._stager is not None:
            self._stager.close()
            self._stager = None
        if self.shared_memory:
            self.shared_memory.close()

    def unlink(self):
        if not self.shared_memory:
            self.init_shared_memory()  # may have been created by another process
        if self.shared_memory:
            self.shared_memory.unlink()
        if self.metadata:
            self.metadata.unlink()

    def reset(self):
        self._need_creation = True

    def init_shared_memory(self, create=False, size=0):
        self.shared_memory = _create_shared_memory(self._shm_name, create=create, size=size)
        self._need_creation = False

    def _create_tensor_meta(self, value):
        """Meta of one leaf at the current end of the layout (API parity with
        the reference's planner hook, ckpt_saver.py:286-301)."""
        if not torch.is_tensor(value):
            return value
        meta = TensorMeta(shape=tuple(value.shape), dtype=value.dtype,
                          element_size=value.element_size(), numel=value.numel(),
                          offset=self._buffer_size)
        self._buffer_size += value.numel() * value.element_size()
        return meta

    def _finish_at_exit(self):
        try:
            self.wait_pending(timeout=120)
        except BaseException:
            pass

    # -- pending drain ----------------------------------------------------------------
    def pending_save(self) -> Optional[PendingSave]:
        p = self._pending
        return p if p is not None and not p.done() else None

    def wait_pending(self, timeout: Optional[float] = None) -> bool:
        p = self._pending
        if p is None:
            return True
        try:
            ok = p.wait(timeout)
        except BaseException:
            self._pending = None  # a failed drain is reported once
            raise
        if ok:
            self._pending = None
        return ok

    # -- save ---------------------------------------------------------------------------
    def _stager_for(self, tensors) -> _DeviceStager:
        dev = tensors[0].device
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        for t in tensors:
            if t.device != dev:
                raise ValueError(
                    "all CUDA tensors of one checkpoint shard must live on one device; "
                    f"got {dev} and {t.device}")
        if self._stager is None or self._stager.device_index != index:
            if self._stager is not None:
                self._stager.close()
            self._stager = _DeviceStager(index)
        return self._stager

    def ensure_segment(self, total: int):
        """Map a segment of exactly `total` bytes (re-creating it on a size
        change)."""
        if self.shared_memory is not None and self.shared_memory.size == total \
                and not self._need_creation:
            return
        if self._stager is not None:
            self._stager.detach()
        if self.shared_memory is not None:
            self.shared_memory.close()
            self.shared_memory = None
        self.init_shared_memory(create=True, size=total)
        self._buffer_size = total

    def write_ranges(self, device_ranges, host_ranges, raw_chunks=(), *, blocking=True,
                     stream=None, finish: Optional[Callable[[], None]] = None,
                     keepalive: Optional[list] = None,
                     pre_drain: Optional[Callable[[], None]] = None,
                     on_error: Optional[Callable[[], None]] = None,
                     in_place: Optional[bool] = None):
        """Move bytes into the (already sized) segment.

        device_ranges / host_ranges: (tensor, segment offset, nbytes) for CUDA /
        CPU tensors; raw_chunks: (bytes-like, offset).  CUDA ranges go through
        one gather kernel + DMA drain, the rest is written by the host right
        away.  `finish` runs once everything has landed — inline when blocking,
        else on the completion thread of the returned PendingSave.
        in_place (default: self.in_place): no snapshot — the drain reads the CUDA
        tensors themselves, so they must not be written until the save is done
        (wait_pending / PendingSave.wait).  Also chosen, together with blocking,
        when HBM has no room for the snapshot arena.
        """
        keepalive = keepalive if keepalive is not None else []
        for chunk, off in raw_chunks:
            view = memoryview(chunk).cast("B")
            self.shared_memory.buf[off:off + view.nbytes] = view
        if host_ranges:
            keep, ptrs, offs, lens = [], [], [], []
            for t, off, nbytes in host_ranges:
                c = t.detach()
                if not c.is_contiguous():
                    c = c.contiguous()
                keep.append(c)
                ptrs.append(c.data_ptr())
                offs.append(off)
                lens.append(nbytes)
            native.host_pack(self.shared_memory.address, ptrs, offs, lens, _host_threads())
            del keep
        ctx, ticket = None, 0
        if device_ranges:
            stager = self._stager_for([r[0] for r in device_ranges])
            stager.attach(self.shared_memory)
            current = torch.cuda.current_stream(stager.device_index)
            if stream is None:
                stream = current
            elif stream != current:
                # snapshot on a side stream: it must see everything the training
                # stream has enqueued so far, and the caller must make the training
                # stream wait for `last_pack_event` before it MUTATES the tensors
                stream.wait_stream(current)
            plan = stager.plan_for(device_ranges, keepalive, role="save", stream=stream)
            in_place = self.in_place if in_place is None else in_place
            cut = None  # hybrid: segment offset from which the tensors are snapshotted
            if in_place:
                arena = stager.ARENA_NONE
                cut = self._hybrid_cut(stager, plan, device_ranges)
                if cut is not None and cut <= min(r[1] for r in device_ranges) \
                        and stager.ensure_arena(plan) == stager.ARENA_FULL:
                    in_place, cut, arena = False, None, stager.ARENA_FULL  # everything fits
            else:
                arena = stager.ensure_arena(plan)
            if arena == stager.ARENA_NONE and not in_place:
                # nobody promised to keep the tensors unchanged: drain before returning
                in_place = blocking = True
            # bounded-arena saves drain inside save_async: announce first, inline
            windowed = arena == stager.ARENA_WINDOWED
            hold = pre_drain is not None and not blocking and not windowed
            if pre_drain is not None and not hold:
                pre_drain()
                pre_drain = None
            if in_place and cut is not None:
                ticket = plan.save_hybrid_async(self.shared_memory.address, cut, stream, hold=hold)
            elif in_place:
                ticket = plan.save_direct_async(self.shared_memory.address, stream, hold=hold)
            else:
                ticket = plan.save_async(self.shared_memory.address, stream, hold=hold)
            self._last_ticket = (stager.ctx, ticket)
            self.last_hybrid_cut = cut
            self.last_save_in_place = in_place
            if in_place:
                self.last_pack_event = None  # there is no snapshot to wait for, only the drain
            else:
                ev = torch.cuda.Event()
                ev.record(stream)
                self.last_pack_event = ev
            ctx = stager.ctx
        if ctx is None and pre_drain is not None:
            pre_drain()
            pre_drain = None
        pending = PendingSave(ctx, ticket, finish or (lambda: None), keepalive,
                              pre_drain=pre_drain if ctx is not None else None,
                              on_error=on_error)
        self._pending = pending
        if blocking or ctx is None:
            pending._complete()
            self.last_timings = pending.timings
            self._pending = None
            pending.wait()  # re-raise a drain error
            return None
        threading.Thread(target=self._run_completion, args=(pending,), name="fc-drain",
                         daemon=True).start()
        return pending

    MIN_SNAPSHOT_BYTES = 64 << 20  # below this a snapshot part is not worth a kernel

    def _hybrid_cut(self, stager: _DeviceStager, plan, device_ranges) -> Optional[int]:
        """In-place save with a snapshot budget: the segment offset from which the
        tensors fit into the arena (None: no budget / no arena -> pure in-place)."""
        budget = self.in_place_snapshot_bytes
        if budget <= 0:
            return None
        want = min(budget, plan.arena_end)
        if self._snapshot_ceiling:
            want = min(want, self._snapshot_ceiling)
        while stager.ctx.arena_info()[1] < want:
            try:
                stager.ctx.arena_reserve(want)
            except native.NativeError as e:
                if e.code != native.FC_ENOMEM:
                    raise
                # no room for that much: remember it, try half (a failed grow has
                # released the previous arena to make room)
                self._snapshot_ceiling = want = want // 2
                if want < self.MIN_SNAPSHOT_BYTES:
                    break
        cap = min(stager.ctx.arena_info()[1], budget)
        end, cut = plan.arena_end, None
        for _, off, _ in sorted(device_ranges, key=lambda r: r[1], reverse=True):
            if end - off > cap:
                break
            cut = off
        return cut

    def wait_snapshot(self, stream=None):
        """Order the caller's next WRITE to the saved tensors after the last save
        has read them: with a snapshot (default) `stream` (default: current) waits on
        the GPU for the gather kernel; after an in-place save the host waits for the
        drain.  Cheap no-op when nothing is pending."""
        if self.last_save_in_place:
            if self._last_ticket is not None and self.pending_save() is not None:
                ctx, ticket = self._last_ticket
                try:
                    ctx.save_sources_wait(ticket)  # in-place part drained (+ gather done)
                    return
                except native.NativeError:
                    pass  # a later ticket took over / drain error: fall through
            self.wait_pending()
            return
        ev = self.last_pack_event
        if ev is not None:
            (stream or torch.cuda.current_stream()).wait_event(ev)

    def save_state_dict(self, state_dict, blocking: bool = True, stream=None,
                        on_complete: Optional[Callable[[], None]] = None,
                        on_error: Optional[Callable[[], None]] = None):
        """Serialise `state_dict` into the segment.

        blocking=True (the reference's semantics): returns None after every
        byte is in shared memory and the meta says writing_shm=False.
        blocking=False: returns a PendingSave right after the gather kernel is
        enqueued on `stream` (default: the current stream); a completion thread
        finishes the protocol and then calls `on_complete`.
        """
        self.wait_pending()
        lay = self._layout = plan_layout(state_dict, self._layout)
        if lay.total > 0:
            self.ensure_segment(lay.total)
        meta_dict = lay.meta
        conf: CheckpointConfig = meta_dict[DLROVER_CKPT_CONFIG_KEY]
        conf.writing_shm = True

        def announce():
            report_local_event(EventReportConstants.TYPE_INFO, str(conf.rank),
                               EventReportConstants.ACTION_MEM_CKPT_START, f"step={conf.step}")
            self.metadata.set(meta_dict)

        # Host-resident leaves are written by this thread right now, so the agent
        # must already know the segment is changing.  With device leaves only,
        # nothing touches the segment before the drain: the announcement (a
        # pickle + two socket round trips) moves to the completion thread and
        # the drain is held until it is out.
        defer_announce = (not blocking) and bool(lay.device_leaves) and not lay.host_leaves
        if not defer_announce:
            announce()

        def finish():
            conf.writing_shm = False
            self.metadata.set(meta_dict)
            report_local_event(EventReportConstants.TYPE_INFO, str(conf.rank),
                               EventReportConstants.ACTION_MEM_CKPT_COMPLETE,
                               f"step={conf.step}")
            if on_complete is not None:
                on_complete()

        def triples(leaves):
            return [(t, m.offset, m.numel * m.element_size) for t, m in leaves]

        # the tensors must outlive the gather kernel
        keepalive = [state_dict] if lay.device_leaves else []
        return self.write_ranges(triples(lay.device_leaves), triples(lay.host_leaves),
                                 blocking=blocking, stream=stream, finish=finish,
                                 keepalive=keepalive,
                                 pre_drain=announce if defer_announce else None,
                                 on_error=on_error)

    def _run_completion(self, pending: PendingSave):
        pending._complete()
        self.last_timings = pending.timings

    # -- load ---------------------------------------------------------------------------
    def load_state_dict(self):
        """Returns the state dict (CPU tensors aliasing the segment), or {} when
        there is no usable in-memory checkpoint."""
        self.wait_pending()
        meta_dict = self.metadata.get()
        config = meta_dict.get(DLROVER_CKPT_CONFIG_KEY, CheckpointConfig())
        if not meta_dict or config.writing_shm:
            return {}
        if self.shared_memory is None or self._need_creation:
            self.init_shared_memory(create=False)
        if not self.shared_memory:
            return {}
        report_local_event(EventReportConstants.TYPE_INFO, str(config.rank),
                           EventReportConstants.ACTION_RESUME_MEM_CKPT_START,
                           f"step={config.step}")
        state_dict = _read_state_dict_from_shm(meta_dict, self.shared_memory)
        report_local_event(EventReportConstants.TYPE_INFO, str(config.rank),
                           EventReportConstants.ACTION_RESUME_MEM_CKPT_COMPLETE,
                           f"step={config.step}")
        return state_dict

    def restore_into(self, target, stream=None, strict: bool = True) -> Dict[str, float]:
        """Scatter the in-memory checkpoint straight into the live tensors of
        `target` (same tree structure as what was saved; extra keys on either
        side raise when strict).  CUDA leaves: one DMA fill of the arena + one
        scatter kernel; CPU leaves: copied from the segment.  Non-tensor leaves
        of the checkpoint are returned under "extras" untouched.

        Replaces `sd = load_state_dict(); model.load_state_dict(sd)` (per
        parameter H2D copies from pageable memory, reference ckpt_saver.py:
        144-161 + user code)."""
        self.wait_pending()
        meta_dict = self.metadata.get()
        config = meta_dict.get(DLROVER_CKPT_CONFIG_KEY, CheckpointConfig())
        if not meta_dict or config.writing_shm:
            raise RuntimeError("no consistent in-memory checkpoint to restore from")
        if self.shared_memory is None or self._need_creation:
            self.init_shared_memory(create=False)
        if not self.shared_memory:
            raise RuntimeError("the checkpoint segment does not exist")

        device_pairs: List[Tuple[torch.Tensor, TensorMeta]] = []
        host_pairs: List[Tuple[torch.Tensor, TensorMeta]] = []

        def walk(tgt, meta, path):
            if isinstance(meta, Mapping):
                if not isinstance(tgt, Mapping):
                    raise KeyError(f"{path}: checkpoint has a dict, target has {type(tgt).__name__}")
                for k, m in meta.items():
                    if k == DLROVER_CKPT_CONFIG_KEY:
                        continue
                    if k not in tgt:
                        if strict:
                            raise KeyError(f"{path}{k}: missing in target")
                        continue
                    walk(tgt[k], m, f"{path}{k}.")
                if strict:
                    extra = [k for k in tgt if k not in meta]
                    if extra:
                        raise KeyError(f"{path}: target keys not in checkpoint: {extra}")
            elif isinstance(meta, list):
                if not isinstance(tgt, list) or len(tgt) != len(meta):
                    raise KeyError(f"{path}: list length/type mismatch")
                for i, m in enumerate(meta):
                    walk(tgt[i], m, f"{path}{i}.")
            elif isinstance(meta, TensorMeta):
                if not torch.is_tensor(tgt):
                    raise KeyError(f"{path}: checkpoint has a tensor, target has {type(tgt).__name__}")
                if tuple(tgt.shape) != tuple(meta.shape) or tgt.dtype != meta.dtype:
                    raise ValueError(
                        f"{path}: shape/dtype mismatch {tuple(tgt.shape)}/{tgt.dtype} vs "
                        f"{tuple(meta.shape)}/{meta.dtype}")
                if meta.numel:
                    (device_pairs if tgt.is_cuda else host_pairs).append((tgt, meta))

        walk(target, meta_dict, "")
        stats = {"device_bytes": 0.0, "host_bytes": 0.0, "fill_ms": 0.0, "scatter_ms": 0.0}
        with torch.no_grad():
            strided = [(t, m) for t, m in host_pairs if not t.is_contiguous()]
            dense = [(t, m) for t, m in host_pairs if t.is_contiguous()]
            for t, m in strided:  # rare: let torch handle the strides
                t.copy_(_read_tensor_from_buf(m, self.shared_memory))
            if dense:
                native.host_unpack(self.shared_memory.address, [t.data_ptr() for t, _ in dense],
                                   [m.offset for _, m in dense],
                                   [m.numel * m.element_size for _, m in dense], _host_threads())
            stats["host_bytes"] = float(sum(m.numel * m.element_size for _, m in host_pairs))
            if device_pairs:
                for t, _ in device_pairs:
                    if not t.is_contiguous():
                        raise ValueError("restore_into needs contiguous CUDA targets")
                stager = self._stager_for([t for t, _ in device_pairs])
                stager.attach(self.shared_memory)
                if stream is None:
                    stream = torch.cuda.current_stream(stager.device_index)
                plan = stager.plan_for(
                    [(t, m.offset, m.numel * m.element_size) for t, m in device_pairs], [],
                    role="restore", stream=stream)
                stats.update(self._run_restore(stager, plan, stream))
        return stats

    DIRECT_RESTORE_MIN_SPAN = 1 << 20  # average span size from which in-place restore wins

    def _run_restore(self, stager: _DeviceStager, plan, stream) -> Dict[str, float]:
        """Few large spans: H2D DMA straight into the targets.  Many small ones: one
        DMA per merged segment run into the arena + one scatter kernel.
        DLROVER_B200_RESTORE=direct|arena overrides the choice."""
        direct = plan.payload_bytes >= plan.n_spans * self.DIRECT_RESTORE_MIN_SPAN
        forced = os.getenv("DLROVER_B200_RESTORE", "")
        if forced in ("direct", "arena"):
            direct = forced == "direct"
        if not direct and stager.ensure_arena(plan) == stager.ARENA_NONE:
            direct = True
        plan.restore_async(self.shared_memory.address, stream, direct=direct)
        stager.ctx.restore_wait()
        fill, scatter, _ = stager.ctx.restore_timings()
        self.last_restore_stats = {"device_bytes": float(plan.payload_bytes), "fill_ms": fill,
                                   "scatter_ms": scatter, "direct": float(direct)}
        return dict(self.last_restore_stats)

    def read_ranges(self, device_ranges, stream=None) -> Dict[str, float]:
        """Inverse of write_ranges for CUDA targets: (tensor, segment offset,
        nbytes) triples are filled by one DMA of the covered segment runs into
        the arena + one scatter kernel into the (contiguous) tensors."""
        self.wait_pending()
        if not device_ranges:
            return {"device_bytes": 0.0, "fill_ms": 0.0, "scatter_ms": 0.0}
        if self.shared_memory is None or self._need_creation:
            self.init_shared_memory(create=False)
        if not self.shared_memory:
            raise RuntimeError("the checkpoint segment does not exist")
        for t, off, nbytes in device_ranges:
            if not (t.is_cuda and t.is_contiguous()):
                raise ValueError("read_ranges needs contiguous CUDA targets")
            if off + nbytes > self.shared_memory.size or t.numel() * t.element_size() != nbytes:
                raise ValueError("read_ranges: range does not fit the segment / the tensor")
        stager = self._stager_for([r[0] for r in device_ranges])
        stager.attach(self.shared_memory)
        if stream is None:
            stream = torch.cuda.current_stream(stager.device_index)
        plan = stager.plan_for(device_ranges, [], role="restore", stream=stream)
        return self._run_restore(stager, plan, stream)

    # -- queries ------------------------------------------------------------------------
    def no_checkpoint_state(self):
        """True when the meta dict holds no config or step 0.  (The agent-side
        handler maps the segment lazily, so the mapping itself says nothing.)"""
        config = self.metadata.get().get(DLROVER_CKPT_CONFIG_KEY, None)
        return config is None or config.step == 0

    def get_checkpoint_config(self, default_config):
        return self.metadata.get().get(DLROVER_CKPT_CONFIG_KEY, default_config)

    def get_master_client(self):
        return self._master_client

    def set_master_client(self, client):
        self._master_client = client
