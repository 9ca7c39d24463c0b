"""Node-local IPC between training processes and the agent that hosts the
checkpoint savers: a lock, a queue and a dict served over unix sockets, and a
POSIX shared-memory segment that outlives the process that created it.

Interface and wire contract follow dlrover/python/common/multi_process.py
(reference @ 468d632):
  * socket files  /tmp/ckpt_sock/[<TORCHELASTIC_RUN_ID>/]<classname>_<name>.sock
    (:32, :216-225);
  * framing: 4-byte big-endian length + pickle of a SocketRequest / *Response
    dataclass (:106-177);
  * SharedLock  (:263-416): persistent client connection, the server drops a
    lock held by a connection that dies (:331-333);
  * SharedQueue (:455-564), SharedDict (:579-672) with its companion queue
    "shard_dict_<name>" that makes a server-side get() wait for in-flight sets;
  * SharedMemory (:675-747): shm_open + mmap WITHOUT Python's resource_tracker,
    so a crashing trainer does not unlink the checkpoint.

The implementation is new: one generic request/response server with a
per-class method table and a thread per connection (the reference serialises
queue/dict requests on the accept thread), exact-length framed reads, and a
client-side mutex so the drain-completion thread and the training thread can
share one lock connection.
"""

from __future__ import annotations

import ctypes
import mmap
import os
import pickle
import queue
import shutil
import socket
import threading
import time
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import _posixshmem

from .constants import NodeEnv
from .log import default_logger as logger

SOCKET_TMP_DIR = "/tmp/ckpt_sock/"

SUCCESS_CODE = "OK"
ERROR_CODE = "ERROR"


# --------------------------------------------------------------- wire messages --


@dataclass
class SocketRequest:
    method: str = ""
    id: str = ""
    args: Dict[str, object] = field(default_factory=dict)


@dataclass
class SocketResponse:
    status: str = ""


@dataclass
class LockAcquireResponse(SocketResponse):
    acquired: bool = False


@dataclass
class LockedResponse(SocketResponse):
    locked: bool = False


@dataclass
class QueueGetResponse(SocketResponse):
    obj: object = None


@dataclass
class QueueSizeResponse(SocketResponse):
    size: int = 0


@dataclass
class QueueEmptyResponse(SocketResponse):
    empty: bool = False


@dataclass
class DictMessage(SocketResponse):
    meta_dict: object = None


# --------------------------------------------------------------------- framing --


def _send_frame(sock: socket.socket, payload: bytes) -> None:
    sock.sendall(len(payload).to_bytes(4, "big") + payload)


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    chunks = []
    while n > 0:
        part = sock.recv(min(n, 1 << 20))
        if not part:
            raise EOFError("peer closed the connection")
        chunks.append(part)
        n -= len(part)
    return b"".join(chunks)


def _recv_frame(sock: socket.socket) -> bytes:
    size = int.from_bytes(_recv_exact(sock, 4), "big")
    return _recv_exact(sock, size) if size else b""


def retry_socket(func):
    """Decorator: retry `func` once a second while the peer's socket is not
    there yet (kwarg ``retry``, default 30), then let the error through."""

    def wrapper(self, *args, **kwargs):
        for _ in range(kwargs.get("retry", 30)):
            try:
                return func(self, *args, **kwargs)
            except (FileNotFoundError, ConnectionRefusedError):
                time.sleep(1)
        return func(self, *args, **kwargs)

    return wrapper


class LockState(int):
    """Result of a client-side SharedLock.locked(): truthy iff locked, and also
    answers `.locked` (the reference hands back the response object there)."""

    @property
    def locked(self) -> bool:
        return bool(self)


def clear_sock_dir():
    shutil.rmtree(SOCKET_TMP_DIR, ignore_errors=True)


def _socket_root() -> str:
    run_id = os.getenv(NodeEnv.TORCHELASTIC_RUN_ID, "")
    return os.path.join(SOCKET_TMP_DIR, run_id) if run_id else SOCKET_TMP_DIR


# ------------------------------------------------------------------ base class --


class LocalSocketComm:
    """A named object shared between the processes of one node.

    ``create=True`` makes this instance the owner: it binds the unix socket and
    answers requests.  ``create=False`` makes it a client of the owner living in
    another process.  Subclasses fill ``_methods`` (name -> callable returning a
    response dataclass) and may set ``_persistent`` to keep one connection per
    client.
    """

    _persistent = False

    def __init__(self, name: str = "", create: bool = False, persist: Optional[bool] = None):
        self._name = name
        self._create = create
        self._persist = self._persistent if persist is None else persist
        self._socket_file = self._socket_path()
        self._server: Optional[socket.socket] = None
        self._client: Optional[socket.socket] = None
        self._client_mutex = threading.Lock()
        if create:
            self._start_server()

    # -- identity -----------------------------------------------------------------
    @property
    def name(self) -> str:
        return self._name

    def _socket_path(self) -> str:
        root = _socket_root()
        os.makedirs(root, exist_ok=True)
        return os.path.join(root, f"{type(self).__name__.lower()}_{self._name}.sock")

    def is_available(self) -> bool:
        try:
            return os.path.exists(self._socket_file)
        except OSError:
            return False

    def unlink(self):
        try:
            os.unlink(self._socket_file)
        except FileNotFoundError:
            pass

    # -- server side ----------------------------------------------------------------
    def _start_server(self):
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        try:
            if os.path.exists(self._socket_file):
                os.unlink(self._socket_file)
            srv.bind(self._socket_file)
            srv.listen(64)
        except OSError:
            srv.close()
            logger.error(f"cannot serve {self._socket_file}", exc_info=True)
            raise
        self._server = srv
        threading.Thread(target=self._accept_loop, name=f"ipc-{self._name}", daemon=True).start()

    def _accept_loop(self):
        srv = self._server
        while srv is not None and srv.fileno() >= 0:
            try:
                conn, _ = srv.accept()
            except OSError:
                return  # server socket closed
            threading.Thread(target=self._serve, args=(conn,), daemon=True).start()

    def _serve(self, conn: socket.socket):
        """Answer requests on one connection until the peer goes away."""
        held = {"lock": False}
        try:
            while True:
                try:
                    request: SocketRequest = pickle.loads(_recv_frame(conn))
                except (EOFError, ConnectionError, OSError):
                    break
                except Exception as e:  # undecodable request: connection is fine
                    logger.error(f"{type(self).__name__}({self._name}): bad request: {e}")
                    _send_frame(conn, pickle.dumps(SocketResponse(status=ERROR_CODE)))
                    continue
                try:
                    response = self._dispatch(request, held)
                    response.status = SUCCESS_CODE
                except Exception as e:
                    logger.error(f"{type(self).__name__}({self._name}).{request.method}: {e}")
                    response = SocketResponse(status=ERROR_CODE)
                finally:
                    self._after_request(request)
                try:
                    _send_frame(conn, pickle.dumps(response))
                except (ConnectionError, OSError):
                    break
                # keep serving this connection until the peer closes it: one-shot
                # clients (the reference's queue/dict clients) close right after
                # the reply, persistent ones save a connect + thread per request
        finally:
            self._on_disconnect(held)
            try:
                conn.close()
            except OSError:
                pass

    def _dispatch(self, request: SocketRequest, held: dict) -> SocketResponse:
        raise NotImplementedError

    def _after_request(self, request: SocketRequest):
        pass

    def _on_disconnect(self, held: dict):
        pass

    # -- client side ------------------------------------------------------------------
    def _request(self, request: SocketRequest, retry: int = 30):
        """Send one request; while the owner's socket does not exist yet (the
        agent may still be starting) retry once a second, `retry` times."""
        payload = pickle.dumps(request)
        with self._client_mutex:
            attempt = 0
            stale_retry = self._persist and self._client is not None
            while True:
                try:
                    if self._client is None:
                        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                        try:
                            c.connect(self._socket_file)
                        except Exception:
                            c.close()
                            raise
                        self._client = c
                    _send_frame(self._client, payload)
                    reply = _recv_frame(self._client)
                    if not self._persist:
                        self._drop_client()
                    return pickle.loads(reply)
                except (FileNotFoundError, ConnectionRefusedError):
                    self._drop_client()
                    attempt += 1
                    if attempt > retry:
                        raise
                    time.sleep(1)
                except (EOFError, BrokenPipeError, ConnectionResetError):
                    # a kept connection the owner has closed meanwhile (owner
                    # restarted, or it serves one request per connection):
                    # reconnect once
                    self._drop_client()
                    if not stale_retry:
                        raise
                    stale_retry = False
                except Exception:
                    self._drop_client()
                    raise

    def _drop_client(self):
        if self._client is not None:
            try:
                self._client.close()
            except OSError:
                pass
            self._client = None

    def close(self):
        self._drop_client()
        if self._server is not None:
            srv, self._server = self._server, None
            try:
                srv.close()
            except OSError:
                pass


# -------------------------------------------------------------------------- lock --


class SharedLock(LocalSocketComm):
    """A mutex shared by name between node-local processes.

    The owner holds a ``threading.Lock``; clients acquire/release it over one
    persistent connection.  If that connection drops while the client holds the
    lock, the owner releases it — a trainer that dies mid-checkpoint cannot
    wedge the saver (reference multi_process.py:331-333).
    """

    _persistent = True

    def __init__(self, name: str = "", create: bool = False, owner: str = ""):
        self._lock = threading.Lock() if create else None
        self._id = owner
        # bumps every time the lock is freed: a connection's hold is only valid for
        # the epoch it was taken in, so a forced release on the owner side
        # (AsyncCheckpointSaver.release_locks) invalidates it
        self._epoch = 0
        super().__init__(name, create)

    def _dispatch(self, request, held):
        if request.method == "acquire":
            got = self.acquire(**request.args)
            held["lock"] = bool(got)
            held["epoch"] = self._epoch
            return LockAcquireResponse(acquired=bool(got))
        if request.method == "locked":
            return LockedResponse(locked=self.locked())
        if request.method == "release":
            # only the connection that holds the lock may free it: a stray
            # second release must not drop a lock someone else has since taken
            if held["lock"] and held.get("epoch") == self._epoch:
                self.release()
            held["lock"] = False
            return SocketResponse()
        raise ValueError(f"unknown lock method {request.method!r}")

    def _on_disconnect(self, held):
        if held.get("lock") and held.get("epoch") == self._epoch:
            logger.info(f"SharedLock({self._name}): holder disconnected, releasing.")
            self.release()

    def acquire(self, blocking: bool = True) -> bool:
        if self._lock is not None:
            return self._lock.acquire(blocking=blocking)
        try:
            resp = self._request(SocketRequest("acquire", self._id, {"blocking": blocking}))
            return bool(resp.acquired) if resp.status == SUCCESS_CODE else False
        except Exception as e:
            logger.warning(f"SharedLock({self._name}).acquire failed: {e}")
            return False

    def release(self):
        if self._lock is not None:
            if self._lock.locked():
                try:
                    self._lock.release()
                    self._epoch += 1
                except RuntimeError:
                    pass  # lost a race with another releaser
            return
        self._request(SocketRequest("release", self._id, {}))

    def locked(self):
        if self._lock is not None:
            return self._lock.locked()
        resp = self._request(SocketRequest("locked", self._id, {}))
        return LockState(bool(getattr(resp, "locked", False)))


# ------------------------------------------------------------------------- queue --


class SharedQueue(LocalSocketComm):
    """A FIFO shared by name; the owner holds a ``queue.Queue(maxsize)``."""

    _persistent = True  # client keeps one connection (see LocalSocketComm._serve)

    def __init__(self, name: str = "", create: bool = False, maxsize: int = 1):
        self._queue = queue.Queue(maxsize) if create else None
        super().__init__(name, create)

    @property
    def queue(self):
        return self._queue

    def _dispatch(self, request, held):
        m = request.method
        if m == "put":
            self.put(**request.args)
            return SocketResponse()
        if m == "get":
            return QueueGetResponse(obj=self.get(**request.args))
        if m == "qsize":
            return QueueSizeResponse(size=self.qsize())
        if m == "empty":
            return QueueEmptyResponse(empty=self.empty())
        raise ValueError(f"unknown queue method {m!r}")

    def put(self, obj, block: bool = True, timeout=None):
        if self._queue is not None:
            self._queue.put(obj, block=block, timeout=timeout)
        else:
            self._request(SocketRequest("put", "", {"obj": obj, "block": block,
                                                    "timeout": timeout}))

    def get(self, block: bool = True, timeout=None):
        if self._queue is not None:
            return self._queue.get(block=block, timeout=timeout)
        resp = self._request(SocketRequest("get", "", {"block": block, "timeout": timeout}))
        return resp.obj if resp.status == SUCCESS_CODE else None

    def qsize(self) -> int:
        if self._queue is not None:
            return self._queue.qsize()
        resp = self._request(SocketRequest("qsize", "", {}))
        return resp.size if resp.status == SUCCESS_CODE else -1

    def empty(self) -> bool:
        if self._queue is not None:
            return self._queue.empty()
        resp = self._request(SocketRequest("empty", "", {}))
        return resp.empty if resp.status == SUCCESS_CODE else False


# -------------------------------------------------------------------------- dict --


class SharedDict(LocalSocketComm):
    """A dict pushed by a writer process and read by the owner (and others).

    A writer announces every ``set`` on the companion queue ``shard_dict_<name>``
    before sending it; the owner's ``get`` waits until that queue is drained so
    it never returns a dict older than a set already under way
    (reference multi_process.py:594-597, :614-616, :663-666).
    """

    _persistent = True

    def __init__(self, name: str = "", create: bool = False):
        self._dict: Any = {}
        self._pending = SharedQueue(name=f"shard_dict_{name}", create=create)
        super().__init__(name, create)

    def _dispatch(self, request, held):
        if request.method == "set":
            self.set(**request.args)
            return DictMessage()
        if request.method == "get":
            return DictMessage(meta_dict=self.get(**request.args))
        raise ValueError(f"unknown dict method {request.method!r}")

    def _after_request(self, request):
        # one announcement is consumed per handled request that followed one
        if self._pending.queue is not None and not self._pending.empty():
            try:
                self._pending.queue.get_nowait()
            except queue.Empty:
                pass

    def set(self, new_dict):
        self._dict = new_dict
        if self._server is None:
            self._pending.put(1)
            resp = self._request(SocketRequest("set", "", {"new_dict": new_dict}))
            if resp.status == ERROR_CODE:
                raise RuntimeError("Fail to set metadata!")

    def get(self, local: bool = False):
        if local:
            return self._dict
        if self._server is not None:
            while not self._pending.empty():
                time.sleep(0.1)
            return self._dict
        resp = self._request(SocketRequest("get", "", {}))
        if resp.status == SUCCESS_CODE:
            self._dict = resp.meta_dict
        return self._dict

    def unlink(self):
        super().unlink()
        self._pending.unlink()

    def close(self):
        super().close()
        self._pending.close()


# --------------------------------------------------------------- shared memory --


class SharedMemory:
    """A named POSIX shared-memory segment mapped into this process.

    Deliberately NOT registered with ``multiprocessing.resource_tracker``: the
    segment must survive the death of the training process that created it so
    the agent can persist it and a restarted trainer can reload from it.  It is
    destroyed only by an explicit ``unlink()`` (the agent's job).
    """

    _mode = 0o600

    def __init__(self, name: Optional[str] = None, create: bool = False, size: int = 0):
        if size < 0:
            raise ValueError("'size' must be a positive integer")
        if create and size == 0:
            raise ValueError("'size' must be a positive number different from zero")
        if name is None and not create:
            raise ValueError("'name' can only be None if create=True")
        flags = os.O_RDWR | ((os.O_CREAT | os.O_EXCL) if create else 0)
        self._fd = -1
        self._mmap: Optional[mmap.mmap] = None
        self._dma_mmap: Optional[mmap.mmap] = None
        self._buf: Optional[memoryview] = None
        if name is None:
            while True:
                candidate = "/fc_" + os.urandom(6).hex()
                try:
                    self._fd = _posixshmem.shm_open(candidate, flags, mode=self._mode)
                    self._name = candidate
                    break
                except FileExistsError:
                    continue
        else:
            self._name = name if name.startswith("/") else "/" + name
            self._fd = _posixshmem.shm_open(self._name, flags, mode=self._mode)
        try:
            if create:
                os.ftruncate(self._fd, size)
            self._size = os.fstat(self._fd).st_size
            self._mmap = mmap.mmap(self._fd, self._size)
        except OSError:
            self.unlink()
            self.close()
            raise
        self._buf = memoryview(self._mmap)

    @property
    def name(self) -> str:
        return self._name[1:] if self._name.startswith("/") else self._name

    @property
    def size(self) -> int:
        return self._size

    @property
    def buf(self) -> memoryview:
        return self._buf

    @property
    def address(self) -> int:
        """Virtual address of byte 0 of the mapping `buf` views."""
        return ctypes.addressof(ctypes.c_char.from_buffer(self._mmap))

    @property
    def dma_address(self) -> int:
        """Virtual address of byte 0 of a SECOND mapping of the same object, used only as
        the DMA target that gets page-locked (cudaHostRegister, slice by slice).  CUDA
        looks at the address to decide whether host memory is pinned, and a copy that
        straddles two registered slices fails — tensors handed to the user view `buf`,
        which stays ordinary pageable memory as far as CUDA is concerned (a
        `param.copy_(view)` then behaves exactly like with the reference)."""
        if self._dma_mmap is None:
            self._dma_mmap = mmap.mmap(self._fd, self._size)
        return ctypes.addressof(ctypes.c_char.from_buffer(self._dma_mmap))

    def stale(self) -> bool:
        """True when the NAME no longer refers to the object this mapping came from
        (the creator unlinked and re-created the segment, e.g. with another size) or
        the object has been resized: a reader must re-attach."""
        if self._fd < 0:
            return True
        try:
            mine = os.fstat(self._fd)
            now = os.stat("/dev/shm" + self._name)
        except OSError:
            return True
        return (mine.st_ino, mine.st_dev) != (now.st_ino, now.st_dev) or now.st_size != self._size

    def close(self):
        if self._buf is not None:
            try:
                self._buf.release()
            except BufferError:
                pass  # tensors/arrays still alias the mapping; leave it mapped
            self._buf = None
        if self._mmap is not None:
            try:
                self._mmap.close()
            except BufferError:
                pass
            self._mmap = None
        if self._dma_mmap is not None:
            try:
                self._dma_mmap.close()
            except BufferError:
                pass
            self._dma_mmap = None
        if self._fd >= 0:
            os.close(self._fd)
            self._fd = -1

    def unlink(self):
        if self._name:
            try:
                _posixshmem.shm_unlink(self._name)
            except FileNotFoundError:
                pass
            logger.info(f"Unlink the shared memory {self._name}")

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
