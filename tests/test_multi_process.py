"""Node-local IPC objects: semantics the checkpoint protocol relies on
(reference behaviours: dlrover/python/tests/test_multi_process.py)."""

import os
import pickle
import socket
import threading
import time

import pytest

from dlrover_b200.common import multi_process as mp
from dlrover_b200.common.multi_process import (
    ERROR_CODE,
    SOCKET_TMP_DIR,
    SharedDict,
    SharedLock,
    SharedMemory,
    SharedQueue,
    SocketRequest,
    SocketResponse,
)


def test_socket_path_and_framing(run_env):
    lock = SharedLock("abc", create=True)
    path = os.path.join(SOCKET_TMP_DIR, run_env, "sharedlock_abc.sock")
    assert os.path.exists(path)
    # raw wire: 4-byte big-endian length + pickle(SocketRequest)
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    c.connect(path)
    payload = pickle.dumps(SocketRequest(method="locked", id="", args={}))
    c.sendall(len(payload).to_bytes(4, "big") + payload)
    head = c.recv(4)
    body = b""
    while len(body) < int.from_bytes(head, "big"):
        body += c.recv(4096)
    resp = pickle.loads(body)
    assert resp.status == "OK" and resp.locked is False
    c.close()
    lock.close()


def test_lock_between_owner_and_clients(run_env):
    owner = SharedLock("l", create=True)
    a = SharedLock("l", create=False)
    b = SharedLock("l", create=False)
    assert owner.acquire() and owner.locked()
    assert a.acquire(blocking=False) is False
    assert a.locked().locked and bool(a.locked())
    owner.release()
    assert not owner.locked() and not a.locked()
    assert a.acquire() is True
    assert owner.locked() and b.acquire(blocking=False) is False
    assert owner.acquire(blocking=False) is False
    a.release()
    assert b.acquire() is True
    # releasing a free lock is harmless
    a.release()
    assert owner.locked()
    b.release()
    assert not owner.locked()
    for x in (a, b, owner):
        x.close()


def test_lock_released_when_holder_connection_dies(run_env):
    owner = SharedLock("d", create=True)
    holder = SharedLock("d", create=False)
    bystander = SharedLock("d", create=False)
    assert holder.acquire()
    assert bystander.acquire(blocking=False) is False
    bystander.close()  # a non-holder leaving changes nothing
    time.sleep(0.3)
    assert owner.locked()
    holder.close()  # trainer died while holding the shard lock
    deadline = time.time() + 5
    while owner.locked() and time.time() < deadline:
        time.sleep(0.05)
    assert not owner.locked()
    owner.close()


def test_blocking_acquire_waits_for_release(run_env):
    owner = SharedLock("w", create=True)
    client = SharedLock("w", create=False)
    assert owner.acquire()
    got = {}

    def waiter():
        got["v"] = client.acquire(blocking=True)
        got["t"] = time.time()

    th = threading.Thread(target=waiter)
    t0 = time.time()
    th.start()
    time.sleep(0.5)
    assert "v" not in got
    owner.release()
    th.join(5)
    assert got["v"] is True and got["t"] - t0 >= 0.45
    client.release()
    owner.close()
    client.close()


def test_client_connection_is_shared_between_threads(run_env):
    """The drain-completion thread releases over the same connection the
    training thread acquired on."""
    owner = SharedLock("t", create=True)
    client = SharedLock("t", create=False)
    assert client.acquire()
    th = threading.Thread(target=client.release)
    th.start()
    th.join(5)
    assert not owner.locked()
    errors = []

    def hammer():
        try:
            for _ in range(50):
                client.locked()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=hammer) for _ in range(4)]
    [t.start() for t in ts]
    [t.join(20) for t in ts]
    assert not errors
    owner.close()
    client.close()


def test_queue(run_env):
    owner = SharedQueue("q", create=True)
    client = SharedQueue("q", create=False)
    owner.put(2)
    assert owner.qsize() == 1 and owner.get() == 2
    client.put({"k": [1, 2]})
    assert client.qsize() == 1 and not client.empty()
    assert client.get() == {"k": [1, 2]}
    assert client.empty() and owner.empty()
    assert client.is_available()
    # a blocked put (queue full, maxsize=1) is released by a consumer even
    # though another client talks to the queue meanwhile
    owner.put("first")
    done = []
    th = threading.Thread(target=lambda: (client.put("second"), done.append(1)))
    th.start()
    time.sleep(0.3)
    assert not done and SharedQueue("q", create=False).qsize() == 1
    assert owner.get() == "first"
    th.join(5)
    assert done and owner.get() == "second"
    owner.close()
    client.close()


def test_dict_set_get_and_error(run_env, monkeypatch):
    owner = SharedDict("m", create=True)
    client = SharedDict("m", create=False)
    d = {"a": 1, "b": {"c": [1, 2, 3]}}
    client.set(d)
    d["a"] = 4
    client.set(d)
    assert owner.get() == d
    assert client.get() == d
    assert client.get(local=True) == d
    other = SharedDict("m", create=False)
    assert other.get() == d
    monkeypatch.setattr(client, "_request", lambda *a, **k: SocketResponse(status=ERROR_CODE))
    with pytest.raises(RuntimeError):
        client.set(d)
    owner.unlink()
    owner.close()


def test_big_message_roundtrip(run_env):
    """Meta trees of real models are 100s of KB: framing must not truncate."""
    owner = SharedDict("big", create=True)
    client = SharedDict("big", create=False)
    d = {f"layer.{i}.weight": ("x" * 100, i, [i] * 10) for i in range(5000)}
    client.set(d)
    assert owner.get() == d and SharedDict("big", create=False).get() == d
    owner.close()


def test_request_retries_until_owner_appears(run_env):
    client = SharedQueue("late", create=False)
    owner_box = {}

    def start_owner():
        time.sleep(1.2)
        owner_box["o"] = SharedQueue("late", create=True)

    threading.Thread(target=start_owner).start()
    client.put(7)  # retried once a second until the socket exists
    assert owner_box["o"].get() == 7
    with pytest.raises((FileNotFoundError, ConnectionRefusedError)):
        SharedQueue("never", create=False)._request(SocketRequest("qsize"), retry=1)


def test_retry_decorator():
    class T:
        n = 0

        @mp.retry_socket
        def f(self, retry=30):
            T.n += 1
            raise FileNotFoundError("x")

    with pytest.raises(FileNotFoundError):
        T().f(retry=1)
    assert T.n == 2


def test_shared_memory_survives_close_and_needs_explicit_unlink(run_env):
    name = f"{run_env}_shm"
    with pytest.raises(ValueError):
        SharedMemory(name=name, create=True, size=-1)
    with pytest.raises(ValueError):
        SharedMemory(name=name, create=True, size=0)
    shm = SharedMemory(name=name, create=True, size=4096)
    assert shm.size == 4096 and shm.name == name
    shm.buf[0:4] = b"abcd"
    assert shm.address != 0
    with pytest.raises(FileExistsError):
        SharedMemory(name=name, create=True, size=4096)
    shm.close()
    again = SharedMemory(name=name)  # still there: close() does not unlink
    assert bytes(again.buf[0:4]) == b"abcd" and again.size == 4096
    again.unlink()
    again.close()
    with pytest.raises(FileNotFoundError):
        SharedMemory(name=name, create=False)


def test_shared_memory_not_tracked_by_resource_tracker(run_env):
    """A child that creates the segment and dies must not take it along."""
    import subprocess
    import sys

    name = f"{run_env}_orphan"
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from dlrover_b200.common.multi_process import SharedMemory\n"
        "s = SharedMemory(name=%r, create=True, size=1024); s.buf[0:2] = b'ok'\n"
        "import os; os._exit(1)\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), name)
    )
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DLROVER_LOG_LEVEL="ERROR"))
    time.sleep(0.5)
    s = SharedMemory(name=name)
    assert bytes(s.buf[0:2]) == b"ok"
    s.unlink()
    s.close()


def test_forced_release_invalidates_the_old_holders_claim():
    """AsyncCheckpointSaver.release_locks() frees a lock on the owner side; the client
    that held it must not be able to drop the lock someone else has taken since."""
    import uuid

    from dlrover_b200.common.multi_process import SharedLock

    name = "epoch" + uuid.uuid4().hex[:6]
    owner = SharedLock(name=name, create=True)
    a = SharedLock(name=name, create=False)
    b = SharedLock(name=name, create=False)
    try:
        assert a.acquire(blocking=False)
        owner.release()                      # forced by the owner
        assert b.acquire(blocking=False)     # someone else takes it
        a.release()                          # stale claim: must be ignored
        assert owner.locked()
        a.close()                            # ... also on disconnect
        import time
        time.sleep(0.2)
        assert owner.locked()
        b.release()
        assert not owner.locked()
    finally:
        b.close()
        owner.unlink()
        owner.close()


def test_shared_memory_second_mapping_and_staleness():
    """dma_address is a second mapping of the same object (the one that gets page-locked);
    stale() notices a re-created or resized object behind the name."""
    import ctypes
    import uuid

    from dlrover_b200.common.multi_process import SharedMemory

    name = "twomaps" + uuid.uuid4().hex[:6]
    a = SharedMemory(name=name, create=True, size=8192)
    try:
        assert a.dma_address != a.address and a.dma_address == a.dma_address
        a.buf[100:104] = b"abcd"
        assert ctypes.string_at(a.dma_address + 100, 4) == b"abcd"
        ctypes.memmove(a.dma_address + 200, b"wxyz", 4)
        assert bytes(a.buf[200:204]) == b"wxyz"
        b = SharedMemory(name=name)
        assert not a.stale() and not b.stale()
        a.unlink()                                   # the writer re-creates it with another size
        c = SharedMemory(name=name, create=True, size=4096)
        assert b.stale() and not c.stale()
        b.close()
        c.unlink()
        c.close()
    finally:
        a.close()
