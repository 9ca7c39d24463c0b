"""Storage plug-in behaviours (reference: dlrover/python/tests/test_storage.py)."""

import os

import torch

from dlrover_b200.common.serialize import ClassMeta
from dlrover_b200.common.storage import (
    KeepLatestStepStrategy,
    KeepStepIntervalStrategy,
    PosixDiskStorage,
    PosixStorageWithDeletion,
    get_checkpoint_storage,
)


def test_posix_disk_storage(tmp_path):
    st = PosixDiskStorage()
    p = str(tmp_path / "a" / "x.txt")
    st.safe_makedirs(os.path.dirname(p))
    st.write("100", p)
    assert st.read(p) == "100" and st.exists(p)
    st.write(b"\x00\x01", str(tmp_path / "b.bin"))
    st.write(memoryview(b"\x02\x03"), str(tmp_path / "c.bin"))
    assert open(tmp_path / "c.bin", "rb").read() == b"\x02\x03"
    assert st.read(str(tmp_path / "missing")) == ""
    sd = {"w": torch.arange(4)}
    sp = str(tmp_path / "deep" / "er" / "sd.pt")
    st.write_state_dict(sd, sp, torch.save)
    back = st.read_state_dict(sp, lambda q: torch.load(q, map_location="cpu"))
    assert torch.equal(back["w"], sd["w"])
    assert st.read_state_dict(str(tmp_path / "nope.pt"), torch.load) == {}
    st.safe_move(sp, str(tmp_path / "moved.pt"))
    assert st.exists(str(tmp_path / "moved.pt")) and not st.exists(sp)
    st.safe_remove(str(tmp_path / "moved.pt"))
    st.safe_remove(str(tmp_path / "moved.pt"))
    assert sorted(st.listdir(str(tmp_path))) == ["a", "b.bin", "c.bin", "deep"]
    st.safe_rmtree(str(tmp_path / "deep"))
    st.safe_rmtree(str(tmp_path / "deep"))
    st.commit(1, True)
    meta = st.get_class_meta()
    assert isinstance(meta.instantiate(), PosixDiskStorage)


def _steps(tmp_path, storage, steps):
    tracker = str(tmp_path / "dlrover_latest.txt")
    for s in steps:
        os.makedirs(tmp_path / str(s), exist_ok=True)
        storage.write(str(s), tracker)
        storage.commit(s, True)
    return sorted(int(d) for d in os.listdir(tmp_path) if d.isdigit())


def test_keep_latest(tmp_path):
    st = get_checkpoint_storage(KeepLatestStepStrategy(max_to_keep=2, checkpoint_dir=str(tmp_path)))
    assert isinstance(st, PosixStorageWithDeletion)
    assert _steps(tmp_path, st, [10, 20, 30, 40]) == [30, 40]


def test_keep_interval(tmp_path):
    st = PosixStorageWithDeletion("dlrover_latest.txt",
                                  KeepStepIntervalStrategy(keep_interval=100,
                                                           checkpoint_dir=str(tmp_path)))
    assert _steps(tmp_path, st, [50, 100, 150, 200, 250]) == [100, 200, 250]
    # a failed commit deletes nothing
    st.write("300", str(tmp_path / "dlrover_latest.txt"))
    st.commit(300, False)
    assert os.path.exists(tmp_path / "250")


def test_class_meta_round_trip(tmp_path):
    strat = KeepLatestStepStrategy(3, str(tmp_path))
    st = PosixStorageWithDeletion("dlrover_latest.txt", strat)
    meta = st.get_class_meta()
    assert isinstance(meta, ClassMeta) and meta.class_name == "PosixStorageWithDeletion"
    clone = meta.instantiate()
    assert isinstance(clone, PosixStorageWithDeletion) and clone._tracker_file == "dlrover_latest.txt"
    assert isinstance(get_checkpoint_storage(None), PosixDiskStorage)


def test_parallel_write_is_byte_identical(tmp_path, monkeypatch):
    import numpy as np

    st = PosixDiskStorage()
    monkeypatch.setattr(PosixDiskStorage, "PARALLEL_WRITE_MIN", 1 << 20)
    monkeypatch.setattr(PosixDiskStorage, "PARALLEL_WRITE_THREADS", 4)
    data = np.random.default_rng(0).integers(0, 256, size=(70 << 20) + 12345, dtype=np.uint8)
    p = str(tmp_path / "seg.distcp")
    st.write(memoryview(data), p)
    assert os.path.getsize(p) == data.size
    assert np.array_equal(np.fromfile(p, dtype=np.uint8), data)
    st.write(memoryview(data[:100]), p)  # small again: plain path, truncates
    assert os.path.getsize(p) == 100
