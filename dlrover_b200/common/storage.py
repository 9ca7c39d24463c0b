"""Storage plug-in API used by the agent-side savers and by engine fallbacks.

Interface: dlrover/python/common/storage.py (reference @ 468d632) —
CheckpointStorage (:24-125), PosixDiskStorage (:128-206), deletion strategies
(:209-261), PosixStorageWithDeletion (:264-323), get_checkpoint_storage (:326).
Users subclass CheckpointStorage and hand the class to the agent by name
through get_class_meta() (the agent re-instantiates it, ckpt_saver.py:447-449).
"""

from __future__ import annotations

import os
import shutil
from abc import ABCMeta, abstractmethod
from typing import Callable, List

from .constants import CheckpointConstant
from .log import default_logger as logger
from .serialize import ClassMeta


class CheckpointStorage(metaclass=ABCMeta):
    """Everything the checkpoint path needs from a storage backend."""

    @abstractmethod
    def write(self, content, path):
        """Write str / bytes / memoryview `content` to `path`."""

    @abstractmethod
    def write_state_dict(self, state_dict, path, write_func):
        """Persist `state_dict` at `path` with `write_func(state_dict, path)`."""

    @abstractmethod
    def read(self, path):
        """Return the text content of `path` ("" if it does not exist)."""

    @abstractmethod
    def read_state_dict(self, path, read_func):
        """Return `read_func(path)` ({} if the path does not exist)."""

    @abstractmethod
    def safe_rmtree(self, dir):
        pass

    @abstractmethod
    def safe_remove(self, path):
        pass

    @abstractmethod
    def safe_makedirs(self, dir):
        pass

    @abstractmethod
    def safe_move(self, src_path, dst_path):
        pass

    @abstractmethod
    def commit(self, step: int, success: bool):
        """Called once per step after all shards were (or failed to be) written."""

    @abstractmethod
    def exists(self, path: str):
        pass

    @abstractmethod
    def listdir(self, path: str):
        pass

    @abstractmethod
    def get_class_meta(self):
        """ClassMeta from which another process can rebuild this storage."""


def _parallel_write(path: str, view: memoryview, threads: int, piece: int = 64 << 20):
    """Write a large buffer with `threads` concurrent pwrite() streams (each call
    releases the GIL); same bytes, same file, fsync'd once at the end."""
    from concurrent.futures import ThreadPoolExecutor

    from . import direct_io

    view = view.cast("B")
    total = view.nbytes
    if direct_io.enabled_for(path):
        # block-device backed directory: whole 4 KiB blocks bypass the page cache
        direct_io.write_buffer(path, view, threads)
        return
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        os.ftruncate(fd, total)

        def put(off):
            end = min(total, off + piece)
            while off < end:
                off += os.pwrite(fd, view[off:end], off)

        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(put, range(0, total, piece)))
        os.fsync(fd)
    finally:
        os.close(fd)


class PosixDiskStorage(CheckpointStorage):
    """Local / NFS-like file system; every write is fsync'd before returning.
    Buffers >= 256 MiB (the raw shm segment of the FSDP/DCP saver) are written
    by several pwrite streams."""

    PARALLEL_WRITE_MIN = 256 << 20
    PARALLEL_WRITE_THREADS = max(1, min(8, (os.cpu_count() or 1) // 2))

    def write(self, content, path):
        binary = isinstance(content, (bytes, bytearray, memoryview))
        try:
            if binary and self.PARALLEL_WRITE_THREADS > 1 and \
                    memoryview(content).nbytes >= self.PARALLEL_WRITE_MIN:
                _parallel_write(str(path), memoryview(content), self.PARALLEL_WRITE_THREADS)
                return
            with open(path, "wb" if binary else "w") as f:
                f.write(content)
                f.flush()
                os.fsync(f.fileno())
        except OSError:
            logger.error(f"Failed to write {path} (binary={binary})")
            raise

    def write_state_dict(self, state_dict, path, write_func=None):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        if write_func:
            write_func(state_dict, path)

    def read(self, path, mode="r"):
        if not os.path.exists(path):
            return ""
        with open(path, mode) as f:
            return f.read()

    def read_state_dict(self, path, read_func):
        if not read_func or not os.path.exists(path):
            return {}
        return read_func(path)

    def safe_rmtree(self, dir):
        if os.path.exists(dir):
            shutil.rmtree(dir, ignore_errors=True)

    def safe_remove(self, path):
        if os.path.exists(path):
            os.remove(path)

    def safe_makedirs(self, dir):
        os.makedirs(dir, exist_ok=True)

    def safe_move(self, src_path, dst_path):
        if os.path.exists(src_path) and not os.path.exists(dst_path):
            shutil.move(src_path, dst_path)

    def commit(self, step, success):
        logger.info(f"Succeed {success} in persisting the checkpoint of step {step}.")

    def exists(self, path: str):
        return os.path.exists(path)

    def listdir(self, path: str):
        return os.listdir(path)

    def get_class_meta(self):
        return ClassMeta(module_path=type(self).__module__, class_name=type(self).__name__)


class CheckpointDeletionStrategy(metaclass=ABCMeta):
    @abstractmethod
    def clean_up(self, step: int, delete_func: Callable):
        """Decide whether the checkpoint of `step` goes; delete with
        `delete_func(directory)`."""


class KeepStepIntervalStrategy(CheckpointDeletionStrategy):
    """Keep only steps that are multiples of `keep_interval`."""

    def __init__(self, keep_interval: int, checkpoint_dir: str):
        self._keep_interval = keep_interval
        self._checkpoint_dir = checkpoint_dir

    def clean_up(self, step, delete_func):
        if step % self._keep_interval == 0:
            return
        victim = os.path.join(self._checkpoint_dir, str(step))
        try:
            logger.info(f"Clean path {victim}")
            delete_func(victim)
        except Exception:
            logger.warning(f"Fail to clean path {victim}!")


class KeepLatestStepStrategy(CheckpointDeletionStrategy):
    """Keep a sliding window of the newest `max_to_keep` steps."""

    def __init__(self, max_to_keep: int, checkpoint_dir: str):
        self._max_to_keep = max(max_to_keep, 1)
        self._checkpoint_dir = checkpoint_dir
        self._steps: List[int] = []

    def clean_up(self, step, delete_func):
        self._steps.append(step)
        if len(self._steps) != self._max_to_keep:
            return
        victim = os.path.join(self._checkpoint_dir, str(self._steps.pop(0)))
        try:
            logger.info(f"Clean path {victim}")
            delete_func(victim)
        except Exception:
            logger.warning(f"Fail to clean path {victim}!")


class PosixStorageWithDeletion(PosixDiskStorage):
    """PosixDiskStorage that, on a successful commit, hands the PREVIOUS tracked
    step to a deletion strategy.  The previous step is learnt by peeking at the
    tracker file right before it is overwritten."""

    def __init__(self, tracker_file: str, deletion_strategy: CheckpointDeletionStrategy):
        super().__init__()
        self._tracker_file = tracker_file
        self._deletion_strategy = deletion_strategy
        self._pre_step = 0

    def write(self, content, path):
        path = str(path)  # may arrive as a pathlib.Path
        if path.endswith(self._tracker_file):
            before = self.read(path)
            if before:
                self._pre_step = int(before)
        super().write(content, path)

    def commit(self, step, success):
        super().commit(step, success)
        if success and self._pre_step not in (0, step):
            self._deletion_strategy.clean_up(self._pre_step, shutil.rmtree)

    def get_class_meta(self):
        return ClassMeta(
            module_path=type(self).__module__,
            class_name=type(self).__name__,
            kwargs={"tracker_file": self._tracker_file,
                    "deletion_strategy": self._deletion_strategy},
        )


def get_checkpoint_storage(deletion_strategy=None):
    if deletion_strategy:
        return PosixStorageWithDeletion(
            tracker_file=CheckpointConstant.TRACER_FILE_NAME,
            deletion_strategy=deletion_strategy,
        )
    return PosixDiskStorage()
