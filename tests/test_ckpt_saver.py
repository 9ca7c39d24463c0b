"""Agent-side savers: factory, event loop, shard persist + done-file commit,
tracker files, breakpoint/SIGTERM save, directory layouts
(reference behaviours: dlrover/python/tests/test_ckpt_saver.py)."""

import os
import pickle
import signal
import threading
import time
from pathlib import Path

import pytest
import torch

from dlrover_b200.ckpt_saver import (
    DLROVER_CKPT_CONFIG_KEY,
    AsyncCheckpointSaver,
    CheckpointConfig,
    CheckpointEvent,
    CheckpointEventType,
    CommonDirCheckpointSaver,
    DdpCheckpointSaver,
    DeepSpeedCheckpointSaver,
    FsdpDcpSaver,
    MegatronCheckpointSaver,
    SharedMemoryHandler,
    TempDirCheckpointSaver,
    _create_shared_memory,
    _traverse_state_dict,
)
from dlrover_b200.common.constants import CheckpointConstant
from dlrover_b200.common.multi_process import SharedMemory, SharedQueue
from dlrover_b200.common.serialize import ClassMeta
from dlrover_b200.common.storage import PosixDiskStorage

MODEL = CheckpointConstant.MODEL_STATES_NAME


class SimpleNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = torch.nn.Linear(64, 32)
        self.fc2 = torch.nn.Linear(32, 10)


class SimpleShardingSaver(TempDirCheckpointSaver):
    """Same shape as the reference test's subclass: TempDir flow + own tracker."""

    def persist_to_storage(self, local_shard_id, ckpt_config):
        sd = self._shm_handlers[local_shard_id].load_state_dict()
        for name, path in ckpt_config.paths.items():
            self.storage.write_state_dict(sd[name], path, torch.save)

    def get_tracker_file(self):
        return os.path.join(self.checkpoint_dir, "tracker.txt")

    def update_tracker_file(self, step):
        self.storage.write(str(step), self.get_tracker_file())


def _storage_meta():
    return PosixDiskStorage().get_class_meta()


def _fill(saver, shard, step, path, sd=None):
    sd = sd if sd is not None else {"model": SimpleNet().state_dict(), "step": step}
    full = {MODEL: sd, DLROVER_CKPT_CONFIG_KEY: CheckpointConfig(rank=shard, step=step,
                                                                paths={MODEL: path})}
    saver._shm_handlers[shard].save_state_dict(full)
    return sd


def test_handler_helpers(run_env):
    h = SharedMemoryHandler(1, host=True)
    m = h._create_tensor_meta(torch.rand(10, 10))
    assert (m.numel, m.element_size, m.offset, m.shape, m.dtype) == (100, 4, 0, (10, 10),
                                                                      torch.float32)
    assert h._create_tensor_meta("leaf") == "leaf"
    assert h.load_state_dict() == {}
    h.metadata.set({"step": 100})
    assert h.metadata.get() == {"step": 100}
    sd = {"model": SimpleNet().state_dict(), "step": 100}
    assert _traverse_state_dict(sd, lambda v: v) == sd
    assert _create_shared_memory(f"{run_env}_none", False) is None
    assert _create_shared_memory(f"{run_env}_zero", True, size=0) is None
    a = _create_shared_memory(f"{run_env}_rep", True, size=10240)
    assert a.size == 10240
    b = _create_shared_memory(f"{run_env}_rep", True, size=102400)  # recreated bigger
    assert b.size == 102400
    c = _create_shared_memory(f"{run_env}_rep", True, size=102400)  # same size: attach
    assert c.size == 102400
    b.unlink()
    h.close()


def test_factory_creates_saver_and_tolerates_repeats(run_env, tmp_path):
    AsyncCheckpointSaver.start_async_saving_ckpt()
    sq = SharedQueue(name="factory", create=False)
    meta = ClassMeta(module_path=DdpCheckpointSaver.__module__,
                     class_name=DdpCheckpointSaver.__name__,
                     kwargs={"checkpoint_dir": str(tmp_path), "storage_meta": _storage_meta()})
    sq.put(meta)
    for _ in range(40):
        if AsyncCheckpointSaver.get_ckpt_saver() is not None:
            break
        time.sleep(0.1)
    saver = AsyncCheckpointSaver.get_ckpt_saver()
    assert isinstance(saver, DdpCheckpointSaver)
    AsyncCheckpointSaver.reset()
    assert saver.wait_saving_checkpoint() is False
    # a restarted trainer announces again: same saver, new dir/rank
    meta.kwargs["checkpoint_dir"] = str(tmp_path / "other")
    meta.kwargs["rank"] = 0
    sq.put(meta)
    sq.put(meta)
    time.sleep(0.5)
    assert AsyncCheckpointSaver.get_ckpt_saver() is saver
    assert saver.checkpoint_dir == str(tmp_path / "other")
    # _remove_sub_dir_of_target_path keeps files, drops dirs
    os.makedirs(tmp_path / "td1")
    (tmp_path / "tf1").write_text("x")
    saver._remove_sub_dir_of_target_path(str(tmp_path))
    assert (tmp_path / "tf1").exists() and not (tmp_path / "td1").exists()


def test_close_twice_unlinks_segment(run_env, tmp_path):
    saver = DdpCheckpointSaver(str(tmp_path), _storage_meta())
    saver._shm_handlers[0].init_shared_memory(create=True, size=1024)
    name = saver._shm_handlers[0]._shm_name
    saver.close()
    saver.close()
    with pytest.raises(FileNotFoundError):
        SharedMemory(name=name)


def test_sigterm_saves_memory_and_sigint_cleans(run_env, tmp_path):
    saver = DdpCheckpointSaver(str(tmp_path), _storage_meta())
    sd = _fill(saver, 0, 100, str(tmp_path / "checkpoint.pt"))
    conf = saver._shm_handlers[0].metadata.get()[DLROVER_CKPT_CONFIG_KEY]
    assert conf.writing_shm is False and conf.step == 100
    old_term = signal.signal(signal.SIGTERM, signal.SIG_IGN)
    old_int = signal.getsignal(signal.SIGINT)
    try:
        AsyncCheckpointSaver._saver_instance = saver
        AsyncCheckpointSaver.register_signal_handler()
        signal.getsignal(signal.SIGTERM)(signal.SIGTERM, None)  # saves, then re-raises to SIG_IGN
        assert sorted(os.listdir(tmp_path)) == ["._dlrover_ckpt_stage", "checkpoint.pt",
                                                "dlrover_latest.txt"]
        assert (tmp_path / "dlrover_latest.txt").read_text() == "100"
        back = torch.load(tmp_path / "checkpoint.pt")
        assert back["step"] == 100 and torch.equal(back["model"]["fc1.weight"],
                                                   sd["model"]["fc1.weight"])
        with pytest.raises(KeyboardInterrupt):
            signal.getsignal(signal.SIGINT)(signal.SIGINT, None)
    finally:
        signal.signal(signal.SIGTERM, old_term)
        signal.signal(signal.SIGINT, old_int)
    saver.persist_to_storage(0, None)  # missing config: skipped, no raise


def test_update_shard_event_and_exit(run_env, tmp_path):
    saver = DdpCheckpointSaver(str(tmp_path), _storage_meta())
    th = threading.Thread(target=saver._sync_shm_to_storage, daemon=True)
    th.start()
    saver._shm_handlers[0].init_shared_memory(create=True, size=1024)
    saver._shm_handlers[0].metadata.set({"step": 100})
    saver._event_queue.put(CheckpointEvent(type=CheckpointEventType.UPDATE_SHARD,
                                           global_shard_num=2))
    time.sleep(0.3)
    assert saver.global_shard_num == 2
    assert saver._shm_handlers[0].no_checkpoint_state()
    saver.close()
    th.join(5)
    assert not th.is_alive()


def test_save_event_persists_and_commits(run_env, tmp_path):
    saver = DdpCheckpointSaver(str(tmp_path), _storage_meta())
    th = threading.Thread(target=saver._sync_shm_to_storage, daemon=True)
    th.start()
    _fill(saver, 0, 7, str(tmp_path / "7" / "rank_0.pt"))
    SharedQueue("ckpt_lock_rank_0").put(CheckpointEvent(type=CheckpointEventType.SAVE, step=7))
    for _ in range(100):
        if (tmp_path / "dlrover_latest.txt").exists():
            break
        time.sleep(0.1)
    assert (tmp_path / "dlrover_latest.txt").read_text() == "7"
    assert sorted(os.listdir(tmp_path)) == ["._dlrover_ckpt_stage", "7", "dlrover_latest.txt"]
    assert os.listdir(tmp_path / "._dlrover_ckpt_stage") == []  # done dir cleaned
    assert saver._latest_step == 7 and not saver._any_rank_locked()
    # a SAVE for a step that is not the one in memory is refused
    SharedQueue("ckpt_lock_rank_0").put(CheckpointEvent(type=CheckpointEventType.SAVE, step=9))
    time.sleep(1.5)
    assert (tmp_path / "dlrover_latest.txt").read_text() == "7"
    saver.close()


def test_commit_times_out_without_all_done_files(run_env, tmp_path):
    saver = DdpCheckpointSaver(str(tmp_path), _storage_meta())
    done = tmp_path / ".done" / "10"
    os.makedirs(done)
    saver.global_shard_num = 2
    (done / "0").write_text("done")
    t0 = time.time()
    saver.commit_checkpoint(100, str(done), timeout=2)
    assert time.time() - t0 >= 2 and not done.exists()
    assert not (tmp_path / "dlrover_latest.txt").exists()
    # stop_commit aborts the wait immediately
    os.makedirs(done)
    saver._stop_commit = True
    saver.commit_checkpoint(100, str(done), timeout=60)
    saver.close()


def test_breakpoint_save_rules(run_env, tmp_path):
    saver = DdpCheckpointSaver(str(tmp_path), _storage_meta())
    saver.save_shm_to_storage()  # nothing in memory: skip
    assert os.listdir(tmp_path) == []
    _fill(saver, 0, 100, str(tmp_path / "c.pt"))
    saver._writing_storage = True  # busy: skip
    saver.save_shm_to_storage()
    assert not (tmp_path / "c.pt").exists() and saver._stop_commit is False
    saver._writing_storage = False
    assert saver._shm_locks[0].acquire()  # a trainer died holding the lock: dirty, skip
    saver.save_shm_to_storage()
    assert not (tmp_path / "c.pt").exists()
    saver._shm_locks[0].release()

    class Master:
        def __init__(self, ok):
            self.ok = ok

        def sync_checkpoint(self, step):
            return self.ok

    saver.save_shm_to_storage(timeout=1, master_client=Master(False))  # a node is gone
    assert saver._stop_commit is True and not (tmp_path / "c.pt").exists()
    saver.save_shm_to_storage(timeout=1, master_client=Master(True))
    assert (tmp_path / "c.pt").exists() and (tmp_path / "dlrover_latest.txt").read_text() == "100"
    os.remove(tmp_path / "c.pt")
    saver.save_shm_to_storage()  # already saved: not again
    assert not (tmp_path / "c.pt").exists()
    saver.close()


def test_inconsistent_shard_steps_are_not_saved(run_env, tmp_path):
    saver = CommonDirCheckpointSaver(str(tmp_path), _storage_meta(), local_shard_num=2,
                                     global_shard_num=2)
    _fill(saver, 0, 100, str(tmp_path / "a.pt"))
    _fill(saver, 1, 101, str(tmp_path / "b.pt"))
    saver.save_shm_to_storage()
    assert os.listdir(tmp_path) == []
    assert saver._check_shard_step_consistence(100, timeout=1) is False
    _fill(saver, 1, 100, str(tmp_path / "b.pt"))
    assert saver._check_shard_step_consistence(100, timeout=1) is True
    assert saver.get_latest_start_saving_step() == 100
    saver.save_step_checkpoint(100)
    assert sorted(os.listdir(tmp_path)) == ["._dlrover_ckpt_stage", "a.pt", "b.pt",
                                            "dlrover_latest.txt"]
    assert saver.get_latest_success_save_dir() == (str(tmp_path), 100)
    saver.close()


def test_temp_dir_saver_moves_stage_dir(run_env, tmp_path):
    saver = SimpleShardingSaver(str(tmp_path), _storage_meta())
    final_dir = tmp_path / "checkpoint-100"
    sd = _fill(saver, 0, 100, str(final_dir / "model.pt"))
    saver.save_step_checkpoint(100)
    assert sorted(os.listdir(tmp_path)) == ["._dlrover_ckpt_stage", "checkpoint-100",
                                            "tracker.txt"]
    assert os.listdir(tmp_path / "._dlrover_ckpt_stage") == []
    assert (tmp_path / "tracker.txt").read_text() == "100"
    back = torch.load(final_dir / "model.pt")
    assert torch.equal(back["model"]["fc2.bias"], sd["model"]["fc2.bias"])
    conf = CheckpointConfig(step=1, paths={"a": "/x/1/a.pt", "b": "/y/1/b.pt"})
    with pytest.raises(ValueError):
        saver._replace_path_dir(conf, "/tmp/stage")
    saver.close()


@pytest.mark.parametrize("cls,extra", [(MegatronCheckpointSaver,
                                        "latest_checkpointed_iteration.txt"),
                                       (DeepSpeedCheckpointSaver, "latest")])
def test_framework_tracker_files(run_env, tmp_path, cls, extra):
    saver = cls(str(tmp_path), _storage_meta())
    saver.update_tracker_file(20)
    assert (tmp_path / "dlrover_latest.txt").read_text() == "20"
    assert (tmp_path / extra).read_text() == "20"
    saver.close()


def test_fsdp_dcp_saver_layout(run_env, tmp_path):
    """Raw segment bytes -> __0_0.distcp, pickled metadata -> .metadata
    (reference: test_ckpt_saver.py:440-481)."""
    saver = FsdpDcpSaver(str(tmp_path), _storage_meta())
    handler = saver._shm_handlers[0]
    handler.init_shared_memory(create=True, size=64)
    handler.shared_memory.buf[0:64] = bytes(range(64))
    path = str(tmp_path / "100" / "__0_0.distcp")
    conf = CheckpointConfig(step=100, paths={MODEL: path})
    handler.metadata.set({DLROVER_CKPT_CONFIG_KEY: conf, "dcp_metadata": {"k": "v"},
                          "no_shard_data": {}})
    saver.save_step_checkpoint(100)
    assert sorted(os.listdir(tmp_path)) == ["._dlrover_ckpt_stage", "100", "dlrover_latest.txt"]
    assert sorted(os.listdir(tmp_path / "100")) == [".metadata", "__0_0.distcp"]
    assert (tmp_path / "100" / "__0_0.distcp").read_bytes() == bytes(range(64))
    assert pickle.loads((tmp_path / "100" / ".metadata").read_bytes()) == {"k": "v"}
    assert (tmp_path / "dlrover_latest.txt").read_text() == "100"
    saver.close()


def test_failing_persist_releases_lock_and_reports(run_env, tmp_path):
    class Boom(DdpCheckpointSaver):
        def persist_to_storage(self, local_shard_id, ckpt_config):
            raise IOError("disk full")

    saver = Boom(str(tmp_path), _storage_meta())
    reported = []

    class Master:
        def report_failures(self, payload, level=None):
            reported.append((payload, level))

    saver.set_master_client(Master())
    _fill(saver, 0, 5, str(tmp_path / "x.pt"))
    saver.save_step_checkpoint(5)
    assert not saver._any_rank_locked() and saver._latest_step == 0
    assert saver.wait_saving_checkpoint() is False
    saver._report_failure_to_master("boom")
    assert reported and "boom" in reported[0][0]
    saver.close()


def test_half_written_segment_is_never_committed(run_env, tmp_path):
    """A trainer that dies while filling/draining leaves writing_shm=True and
    (through its dropped connection) a free lock: the agent must neither persist
    nor commit that step."""
    saver = DdpCheckpointSaver(str(tmp_path), _storage_meta())
    _fill(saver, 0, 5, str(tmp_path / "5" / "rank_0.pt"))
    meta = saver._shm_handlers[0].metadata.get()
    meta[DLROVER_CKPT_CONFIG_KEY].writing_shm = True
    saver._shm_handlers[0].metadata.set(meta)
    saver.save_step_checkpoint(5)
    assert not (tmp_path / "dlrover_latest.txt").exists()
    assert not (tmp_path / "5").exists()
    assert saver._latest_step == 0 and not saver._any_rank_locked()
    saver.save_shm_to_storage()  # breakpoint save: same answer
    assert not (tmp_path / "dlrover_latest.txt").exists()
    saver.close()
