"""Parallel writer of `torch.save` files — byte-identical to `torch.save(obj, path)`.

Why: the agent persists each shard with `storage.write_state_dict(sd, path,
torch.save)` (reference ckpt_saver.py:1079-1122).  torch's zip writer copies and
CRC-32s every tensor on ONE thread (measured 0.26 GB/s in the build container,
~1 GB/s on a fast host); the shard lock is held for that long and every memory
checkpoint the trainer attempts meanwhile is skipped.  The payload is already in
host memory (views on the shm segment), so the file can be produced with the
CRCs computed and the bytes written by several threads.

How it stays byte-identical (checked by sha256 in tests/test_fast_torch_save.py):
  * the record list and the pickle (`data.pkl`) come from torch itself:
    `torch.serialization._save` is run against a recording zip object, so names,
    order, storage keys and pickle bytes are whatever this torch version emits;
  * the container is re-stated from what PyTorchStreamWriter/miniz produce:
    local header (flags 0x808, zero crc/sizes) + "FB" padding extra field that
    64-byte-aligns the payload, payload, data descriptor, central directory,
    ZIP64 end records (always present), zip64 extra fields once sizes/offsets
    reach 4 GiB;
  * `version` and the first half of `.data/serialization_id` (a hash over the
    record NAMES that depends on libstdc++'s unordered_set order) are read from
    a throw-away archive written by the real `torch._C.PyTorchFileWriter` with
    the same record names and 1-byte payloads; the second half is
    hash_combine() over the records' CRC-32s in write order.
If anything looks unfamiliar (new record kinds, non-CPU storages) `fast_save`
raises `Unsupported` and callers fall back to `torch.save`.
"""

from __future__ import annotations

import ctypes
import os
import pickle
import struct
import tempfile
import zipfile
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import List, Tuple

import torch
import torch.serialization as _ts

_M64 = (1 << 64) - 1
_U32 = 0xFFFFFFFF
_ALIGN = 64
_FLAGS = 0x0808  # data descriptor + UTF-8 names
_PIECE = 64 << 20


class Unsupported(RuntimeError):
    pass


class _Recorder:
    """Stands in for torch._C.PyTorchFileWriter inside torch.serialization._save."""

    def __init__(self):
        self.records: List[Tuple[str, object, int]] = []
        self.keepalive = []

    def write_record(self, name, data, size):
        if isinstance(data, str):
            data = data.encode()
        if isinstance(data, (bytes, bytearray)):
            self.records.append((name, memoryview(bytes(data[:size])), size))
            return
        if getattr(data, "device", None) is None or data.device.type != "cpu":
            raise Unsupported("only CPU storages can be written by fast_save")
        self.keepalive.append(data)
        if size:
            view = memoryview((ctypes.c_char * size).from_address(data.data_ptr())).cast("B")
        else:
            view = memoryview(b"")
        self.records.append((name, view, size))

    def write_record_metadata(self, name, size):
        raise Unsupported("skip_data serialization is not supported")


def _hash_combine(seed: int, value: int) -> int:
    return (seed ^ ((value + 0x9E3779B9 + ((seed << 6) & _M64) + (seed >> 2)) & _M64)) & _M64


def _surrogate(prefix: str, names: List[str]):
    """(name-hash digits, version payload) from the real writer fed with the
    same record names."""
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, prefix + ".pt")
        w = torch._C.PyTorchFileWriter(path)
        for n in names:
            w.write_record(n, b"x", 1)
        w.write_end_of_file()
        del w
        with zipfile.ZipFile(path) as z:
            listed = [i.filename for i in z.infolist()]
            want = [f"{prefix}/{n}" for n in names] + [f"{prefix}/version",
                                                       f"{prefix}/.data/serialization_id"]
            if listed != want:
                raise Unsupported(f"unexpected record list from torch's writer: {listed[-3:]}")
            sid = z.read(f"{prefix}/.data/serialization_id").decode()
            version = z.read(f"{prefix}/version")
    if len(sid) != 40 or not sid.isdigit():
        raise Unsupported("unexpected serialization_id format")
    return sid[:20], version


def _crc(view: memoryview, threads: int, pool) -> int:
    n = view.nbytes
    if n <= _PIECE or threads <= 1:
        return zlib.crc32(view) & _U32
    # CRC of a concatenation from piecewise CRCs (zlib's crc32_combine restated
    # with GF(2) matrix squaring would do; simpler: chain pieces sequentially per
    # record but run RECORDS in parallel — records are what we have many of).
    c = 0
    for off in range(0, n, _PIECE):
        c = zlib.crc32(view[off:off + _PIECE], c)
    return c & _U32


def _zip64_extra(usize=None, csize=None, offset=None) -> bytes:
    body = b"".join(struct.pack("<Q", v) for v in (usize, csize, offset) if v is not None)
    return struct.pack("<HH", 0x0001, len(body)) + body


def fast_save(obj, path: str, threads: int = 4, pickle_protocol: int = 2) -> None:
    """Write `obj` to `path` exactly as torch.save(obj, path) would."""
    path = os.fspath(path)
    prefix = os.path.splitext(os.path.basename(path))[0]
    rec = _Recorder()
    _ts._save(obj, rec, pickle, pickle_protocol, False)
    records = rec.records
    names = [n for n, _, _ in records]
    if names[:4] != ["data.pkl", ".format_version", ".storage_alignment", "byteorder"] or \
            any(not n.startswith("data/") for n in names[4:]):
        raise Unsupported(f"unexpected records {names[:6]}")
    name_hash, version = _surrogate(prefix, names)

    threads = max(1, threads)
    with ThreadPoolExecutor(max_workers=threads) as pool:
        crcs = list(pool.map(lambda r: _crc(r[1], threads, pool), records))
        combined = 0
        sizes = [r[2] for r in records] + [len(version)]
        for c, n in zip(crcs + [zlib.crc32(version) & _U32], sizes):
            if n:  # empty records do not take part
                combined = _hash_combine(combined, c)
        sid = (name_hash + "%020d" % combined).encode()
        records = records + [("version", memoryview(version), len(version)),
                             (".data/serialization_id", memoryview(sid), len(sid))]
        crcs = crcs + [zlib.crc32(version) & _U32, zlib.crc32(sid) & _U32]

        # ---- layout ---------------------------------------------------------------
        pos = 0
        entries = []  # (name_bytes, crc, size, lfh_offset, data_offset, lfh_bytes, dd_bytes)
        for (name, view, size), crc in zip(records, crcs):
            nm = f"{prefix}/{name}".encode()
            z64 = b""
            big = size >= _U32
            far = pos >= _U32
            if big or far:
                # miniz puts the overflowing values (sizes and/or the header's own
                # offset) into a zip64 extra field; the fixed fields stay zero
                # (the compressed size is not known yet when miniz writes the
                # local header: it stays 0 there)
                z64 = _zip64_extra(size if big else None, 0 if big else None,
                                   pos if far else None)
            fixed = 30 + len(nm) + len(z64) + 4
            pad = (-(pos + fixed)) % _ALIGN
            extra = z64 + b"FB" + struct.pack("<H", pad) + b"Z" * pad
            # empty records carry no data descriptor (and no 0x08 flag)
            flags = _FLAGS if size else _FLAGS & ~0x0008
            lfh = struct.pack("<IHHHHHIIIHH", 0x04034B50, 0, flags, 0, 0, 0, 0, 0, 0, len(nm),
                              len(extra))
            lfh += nm + extra
            data_off = pos + len(lfh)
            if size == 0:
                dd = b""
            elif z64:  # zip64 entry: 8-byte sizes in the data descriptor
                dd = struct.pack("<IIQQ", 0x08074B50, crc, size, size)
            else:
                dd = struct.pack("<IIII", 0x08074B50, crc, size, size)
            entries.append((nm, crc, size, pos, data_off, lfh, dd))
            pos = data_off + size + len(dd)
        cd_offset = pos
        cd = bytearray()
        for nm, crc, size, lfh_off, _, _, _ in entries:
            big = size >= _U32
            far = lfh_off >= _U32
            extra = _zip64_extra(size if big else None, size if big else None,
                                 lfh_off if far else None) if (big or far) else b""
            cd += struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 0, 0,
                              _FLAGS if size else _FLAGS & ~0x0008, 0, 0, 0, crc,
                              _U32 if big else size, _U32 if big else size, len(nm),
                              len(extra), 0, 0, 0, 0, _U32 if far else lfh_off)
            cd += nm + extra
        n = len(entries)
        end = struct.pack("<IQHHIIQQQQ", 0x06064B50, 44, 0x031E, 45, 0, 0, n, n, len(cd),
                          cd_offset)
        end += struct.pack("<IIQI", 0x07064B50, 0, cd_offset + len(cd), 1)
        end += struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, min(n, 0xFFFF), min(n, 0xFFFF),
                           min(len(cd), _U32), min(cd_offset, _U32), 0)
        total = cd_offset + len(cd) + len(end)

        # ---- write ----------------------------------------------------------------
        from .common import direct_io

        if direct_io.enabled_for(path):
            # block-device backed directory: the whole 4 KiB blocks inside each record's
            # payload go through O_DIRECT (page-aligned bounce buffers), headers and
            # partial blocks through the page cache — same bytes at the same offsets
            w = direct_io.DirectWriter(path, total, threads)
            try:
                for (name, view, size), (nm, crc, _, lfh_off, data_off, lfh, dd) in zip(
                        records, entries):
                    w.write_small(lfh, lfh_off)
                    if dd:
                        w.write_small(dd, data_off + size)
                    if size:
                        w.add(view, data_off)
                w.write_small(bytes(cd) + end, cd_offset)
                w.run()
            finally:
                w.close(sync=False)
            del rec
            return
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        try:
            os.ftruncate(fd, total)
            jobs = []
            for (name, view, size), (nm, crc, _, lfh_off, data_off, lfh, dd) in zip(records,
                                                                                   entries):
                os.pwrite(fd, lfh, lfh_off)
                os.pwrite(fd, dd, data_off + size)
                for off in range(0, size, _PIECE):
                    jobs.append((view, off, min(size, off + _PIECE), data_off))
            os.pwrite(fd, bytes(cd) + end, cd_offset)

            def put(job):
                view, off, stop, base = job
                while off < stop:
                    off += os.pwrite(fd, view[off:stop], base + off)

            list(pool.map(put, jobs))
            # no fsync: torch.save does not sync either (the reference only
            # fsyncs its small tracker/done files, storage.py:129-141)
        finally:
            os.close(fd)
    del rec


def save(obj, path, threads: int = 4):
    """fast_save with a torch.save fallback (same bytes either way)."""
    try:
        fast_save(obj, path, threads=threads)
    except Unsupported:
        torch.save(obj, path)
