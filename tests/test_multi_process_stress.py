"""Stress of the socket-served shared objects with several client PROCESSES at once
(the trainer ranks of a node all talk to the agent's servers concurrently):
  * SharedLock: mutual exclusion (a read-modify-write of a file under the lock never
    loses an update) with blocking and non-blocking acquires mixed;
  * SharedQueue: every item put (by 6 processes) is got exactly once (by the owner);
  * SharedDict: set replaces the whole dict (reference semantics); a get concurrent with
    other processes' sets returns ONE writer's whole value, never a mix.
Bounded to a few seconds."""

import multiprocessing as mp
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_PROC, N_OPS = 6, 60


def _client(idx, run_id, counter_file, out_q):
    sys.path.insert(0, ROOT)
    os.environ["TORCHELASTIC_RUN_ID"] = run_id
    os.environ["DLROVER_LOG_LEVEL"] = "ERROR"
    from dlrover_b200.common.multi_process import SharedDict, SharedLock, SharedQueue

    lock = SharedLock("stress", create=False)
    queue = SharedQueue("stress", create=False)
    table = SharedDict("stress", create=False)
    got, failed_try = [], 0
    for i in range(N_OPS):
        if i % 3 == 0:
            while not lock.acquire(blocking=False):  # spin on the non-blocking form
                failed_try += 1
        else:
            assert lock.acquire(blocking=True)
        with open(counter_file, "r+") as f:          # unprotected RMW: only safe under the lock
            v = int(f.read() or 0)
            f.seek(0)
            f.write(str(v + 1))
            f.truncate()
        lock.release()
        queue.put((idx, i))
        table.set({"writer": idx, "n": i + 1, "data": [idx] * (50 * (i + 1))})
        seen = table.get()
        assert seen["data"] == [seen["writer"]] * (50 * seen["n"]), "torn dict"
    out_q.put((idx, got, failed_try))
    for o in (lock, queue, table):
        o.close()


@pytest.mark.timeout(120)
def test_many_processes_hammer_the_shared_objects(run_env, tmp_path):
    from dlrover_b200.common.multi_process import SharedDict, SharedLock, SharedQueue

    lock = SharedLock("stress", create=True)
    queue = SharedQueue("stress", create=True)
    table = SharedDict("stress", create=True)
    counter = tmp_path / "counter"
    counter.write_text("0")
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_client, args=(i, run_env, str(counter), out_q))
             for i in range(N_PROC)]
    try:
        for p in procs:
            p.start()
        # the owner consumes while the clients produce
        seen = [queue.get(timeout=60) for _ in range(N_PROC * N_OPS)]
        results = [out_q.get(timeout=60) for _ in procs]
        for p in procs:
            p.join(30)
            assert p.exitcode == 0
        assert len(results) == N_PROC
        # mutual exclusion: no lost update
        assert int(counter.read_text()) == N_PROC * N_OPS
        assert not lock.locked()
        # queue conservation: every (idx, i) exactly once, nothing left behind
        seen = [tuple(x) for x in seen]
        assert len(seen) == len(set(seen)) == N_PROC * N_OPS
        assert set(seen) == {(p, i) for p in range(N_PROC) for i in range(N_OPS)}
        assert queue.empty()
        final = table.get()
        assert final["n"] == N_OPS and final["data"] == [final["writer"]] * (50 * N_OPS)
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
        for o in (lock, queue, table):
            o.close()
            o.unlink()
