"""The cooperative save's agreement protocol (CheckpointEngine._cooperative_save) with the
device work faked out: who calls the handler, with which save number, who gives up, and that
nobody is ever left waiting.  Real control segment, threads instead of processes."""

import threading
import time

import pytest

from dlrover_b200.ckpt_saver import CheckpointConfig
from dlrover_b200.common.ctl_segment import ControlSegment
from dlrover_b200.flash_checkpoint.engine import FullCheckpointEngine


class _Lock:
    def __init__(self, free=True):
        self.free, self.acquired, self.released = free, 0, 0

    def acquire(self, blocking=True):
        if self.free:
            self.free = False
            self.acquired += 1
            return True
        return False

    def release(self):
        self.free = True
        self.released += 1

    def locked(self):
        return not self.free


class _Handler:
    def __init__(self, ctl, pending=False, boom=False):
        self.ctl, self._pending, self.boom, self.calls = ctl, pending, boom, []

    def pending_save(self):
        return object() if self._pending else None

    def save_state_dict(self, state_dict, blocking=True, on_complete=None, on_error=None,
                        stream=None, coop=None):
        if self.boom:
            raise OSError("segment cannot be created")
        self.calls.append((coop.index, coop.seq, coop.opened))
        if coop.leader:
            coop.ctl.next_coop_seq()          # what the real handler does after announcing
        if on_complete is not None:
            on_complete()


def _engine(ctl, local_rank, monkeypatch, lock=None, handler=None, n=3):
    e = object.__new__(FullCheckpointEngine)
    e._coop_ctl, e._coop_wanted, e._coop_established = ctl, True, False
    e._shm_handler = handler or _Handler(ctl)
    e._shm_lock = lock or _Lock()
    e._local_rank, e._rank, e._group_rank, e._world_size = local_rank, local_rank, 0, n
    e._save_timeout, e._async_drain, e._saver_group = 5, True, None
    e.is_skip, e._cached_step = False, -1
    monkeypatch.setenv("LOCAL_WORLD_SIZE", str(n))
    return e


def _run(engines, step, blocking=False):
    out = [None] * len(engines)

    def call(i, e):
        try:
            out[i] = e._cooperative_save({"w": 1}, CheckpointConfig(step=step, paths={}), blocking)
        except BaseException as err:  # noqa: BLE001
            out[i] = err

    threads = [threading.Thread(target=call, args=(i, e)) for i, e in enumerate(engines)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(20)
    assert not any(t.is_alive() for t in threads), "somebody is still waiting"
    return out


@pytest.fixture
def ctl(run_env):
    c = ControlSegment.create(0)
    yield c
    c.unlink()
    c.close()


def test_everybody_ready(ctl, monkeypatch):
    engines = [_engine(ctl, r, monkeypatch) for r in range(3)]
    for step in (1, 2):
        assert _run(engines, step) == [True, True, True]
    seqs = [e._shm_handler.calls for e in engines]
    assert seqs[0] == [(0, 1, False), (0, 2, False)]                 # the leader opens
    assert seqs[1] == [(1, 1, True), (1, 2, True)] and seqs[2] == [(2, 1, True), (2, 2, True)]
    assert engines[0]._shm_lock.released == 2 and engines[0]._coop_established
    assert all(e._cached_step == 2 for e in engines)


def test_a_follower_with_a_drain_in_flight_makes_everybody_skip(ctl, monkeypatch):
    engines = [_engine(ctl, 0, monkeypatch), _engine(ctl, 1, monkeypatch),
               _engine(ctl, 2, monkeypatch, handler=_Handler(ctl, pending=True))]
    assert _run(engines, 1) == [False, False, False]
    assert all(not e._shm_handler.calls for e in engines)
    assert engines[0]._shm_lock.free and engines[0].is_skip
    engines[2]._shm_handler._pending = False       # next time it works, numbers still in step
    assert _run(engines, 2) == [True, True, True]
    assert engines[1]._shm_handler.calls == [(1, 2, True)]


def test_agent_holds_the_lock(ctl, monkeypatch):
    engines = [_engine(ctl, 0, monkeypatch, lock=_Lock(free=False)), _engine(ctl, 1, monkeypatch),
               _engine(ctl, 2, monkeypatch)]
    assert _run(engines, 1) == [False, False, False]
    assert engines[0]._shm_lock.released == 0      # never acquired, never released


def test_followers_that_never_show_up(ctl, monkeypatch):
    monkeypatch.setenv("DLROVER_B200_COOP_JOIN_TIMEOUT_S", "0.3")
    leader = _engine(ctl, 0, monkeypatch)
    t0 = time.time()
    assert _run([leader], 1) == [FullCheckpointEngine._SOLO]
    assert 0.25 < time.time() - t0 < 5
    assert leader._coop_wanted is False and leader._shm_lock.free
    # a follower that turns up late for that save is told to stay out
    late = _engine(ctl, 1, monkeypatch)
    ctl.slot_arrive(1, ctl.coop_seq(), True)        # it had posted for the save just called off
    assert ctl.wait_coop_open(ctl.coop_seq(), 1) is False


def test_leader_failure_releases_the_followers(ctl, monkeypatch):
    engines = [_engine(ctl, 0, monkeypatch, handler=_Handler(ctl, boom=True)),
               _engine(ctl, 1, monkeypatch), _engine(ctl, 2, monkeypatch)]
    t0 = time.time()
    out = _run(engines, 1)
    assert isinstance(out[0], OSError)
    assert out[1] is False and out[2] is False      # told to stay out, at once
    assert time.time() - t0 < 5
    assert engines[0]._shm_lock.free
    # and the next save works, numbers in step
    engines[0]._shm_handler.boom = False
    assert _run(engines, 2) == [True, True, True]
