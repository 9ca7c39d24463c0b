"""O_DIRECT persist (common/direct_io.py): files are byte-identical to the buffered
writers' — the raw segment dump of the FSDP/DCP saver and the torch.save zip."""

import hashlib
import os

import numpy as np
import pytest
import torch

from dlrover_b200 import fast_torch_save
from dlrover_b200.common import direct_io, storage


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.fixture
def direct_dir(tmp_path, monkeypatch):
    monkeypatch.setenv("DLROVER_B200_DIRECT_IO", "1")
    if not direct_io.enabled_for(str(tmp_path / "x")):
        pytest.skip("this file system refuses O_DIRECT")
    return tmp_path


def test_mode_switch(tmp_path, monkeypatch):
    monkeypatch.setenv("DLROVER_B200_DIRECT_IO", "0")
    assert not direct_io.enabled_for(str(tmp_path / "x"))
    monkeypatch.setenv("DLROVER_B200_DIRECT_IO", "auto")
    assert not direct_io.enabled_for("/dev/shm/whatever")  # tmpfs: pointless
    monkeypatch.delenv("DLROVER_B200_DIRECT_IO")
    assert not direct_io.enabled_for(str(tmp_path / "x"))   # opt-in


@pytest.mark.parametrize("nbytes", [1, 4095, 4096, 4097, (3 << 20) + 123, 40 << 20])
def test_raw_buffer_identical(direct_dir, nbytes):
    rng = np.random.default_rng(nbytes)
    data = rng.integers(0, 256, size=nbytes, dtype=np.uint8)
    view = memoryview(data)
    d, b = direct_io.write_buffer(str(direct_dir / "direct.bin"), view, threads=3)
    assert d + b == nbytes and d == nbytes // 4096 * 4096 if nbytes >= 4096 else d == 0
    assert open(direct_dir / "direct.bin", "rb").read() == data.tobytes()


def test_ranges_at_odd_offsets_and_addresses(direct_dir):
    """Ranges whose file offset and memory address are not congruent mod 4 KiB go through
    the bounce buffer; partial blocks stay buffered; nothing overlaps."""
    rng = np.random.default_rng(5)
    blob = rng.integers(0, 256, size=9 << 20, dtype=np.uint8)
    total = 12 << 20
    want = np.zeros(total, dtype=np.uint8)
    w = direct_io.DirectWriter(str(direct_dir / "odd.bin"), total, threads=4)
    pos, src = 77, 13
    for n in (10, 5000, 4096, 70_001, 1 << 20, (2 << 20) + 7, 3):
        piece = blob[src:src + n]
        w.write_small(b"HDR", pos - 3)
        want[pos - 3:pos] = np.frombuffer(b"HDR", dtype=np.uint8)
        w.add(memoryview(piece), pos)
        want[pos:pos + n] = piece
        pos += n + 1000 + (n % 7)
        src += n + 1
    w.run()
    w.close()
    assert w.direct_bytes > 0 and w.buffered_bytes > 0
    got = np.fromfile(direct_dir / "odd.bin", dtype=np.uint8)
    assert np.array_equal(got, want)


def test_torch_zip_identical_to_torch_save(direct_dir, monkeypatch):
    g = torch.Generator().manual_seed(3)
    sd = {"w": torch.randn(1500, 1501, generator=g), "b": torch.randn(7, generator=g),
          "idx": torch.arange(100_003), "nested": {"h": torch.randn(333, 17).to(torch.bfloat16)},
          "step": 12, "empty": torch.empty(0)}
    (direct_dir / "a").mkdir()
    (direct_dir / "b").mkdir()
    (direct_dir / "c").mkdir()
    torch.save(sd, direct_dir / "a" / "rank_0.pt")
    fast_torch_save.fast_save(sd, str(direct_dir / "b" / "rank_0.pt"), threads=3)   # O_DIRECT
    monkeypatch.setenv("DLROVER_B200_DIRECT_IO", "0")
    fast_torch_save.fast_save(sd, str(direct_dir / "c" / "rank_0.pt"), threads=3)   # buffered
    assert _sha(direct_dir / "a" / "rank_0.pt") == _sha(direct_dir / "b" / "rank_0.pt") \
        == _sha(direct_dir / "c" / "rank_0.pt")
    back = torch.load(direct_dir / "b" / "rank_0.pt")
    assert torch.equal(back["w"], sd["w"]) and back["step"] == 12


def test_storage_write_uses_it_for_big_buffers(direct_dir, monkeypatch):
    monkeypatch.setattr(storage.PosixDiskStorage, "PARALLEL_WRITE_MIN", 1 << 20)
    monkeypatch.setattr(storage.PosixDiskStorage, "PARALLEL_WRITE_THREADS", 3)
    data = np.random.default_rng(1).integers(0, 256, size=(5 << 20) + 17, dtype=np.uint8)
    p = str(direct_dir / "__0_0.distcp")
    storage.PosixDiskStorage().write(memoryview(data), p)
    assert open(p, "rb").read() == data.tobytes()


def test_random_range_sets_with_small_pieces(direct_dir, monkeypatch):
    """Many random layouts (gaps, odd offsets, odd source addresses, ranges smaller and
    larger than a block) with the piece size shrunk to two blocks, so that every split of
    DirectWriter.add / _put is taken: file == the bytes placed at their offsets; the direct
    and buffered byte counts add up to what was queued."""
    monkeypatch.setattr(direct_io, "PIECE", 2 * direct_io.BLOCK)
    rng = np.random.default_rng(11)
    blob = rng.integers(0, 256, size=1 << 20, dtype=np.uint8)
    for case in range(40):
        n_ranges = int(rng.integers(1, 9))
        pos, want_parts = int(rng.integers(0, 9000)), []
        for _ in range(n_ranges):
            n = int(rng.choice([1, 17, 4095, 4096, 4097, 8192, 12_289, 40_000, 100_003]))
            src = int(rng.integers(0, blob.size - n))
            want_parts.append((pos, blob[src:src + n]))
            pos += n + int(rng.choice([0, 1, 4096, 777]))
        total = pos + int(rng.integers(0, 5000))
        path = str(direct_dir / f"r{case}.bin")
        w = direct_io.DirectWriter(path, total, threads=int(rng.integers(1, 5)))
        want = np.zeros(total, dtype=np.uint8)
        for off, piece in want_parts:
            w.add(memoryview(piece), off)
            want[off:off + piece.size] = piece
        w.run()
        w.close()
        assert w.direct_bytes + w.buffered_bytes == sum(p.size for _, p in want_parts)
        assert w.direct_bytes % direct_io.BLOCK == 0
        got = np.fromfile(path, dtype=np.uint8)
        assert np.array_equal(got, want), case
        os.remove(path)
