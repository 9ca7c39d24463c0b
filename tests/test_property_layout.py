"""Property-based parity (hypothesis): random nested state dicts — every dtype width,
0-dim / 0-numel / odd-length tensors, non-contiguous views, lists inside dicts inside
lists, tuples and scalars as leaves — through the product's planner and CPU save path
against the oracle (oracle/shm_layout.py, pinned to the reference's goldens):
identical meta tree, identical segment bytes, identical reload."""

import numpy as np
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from dlrover_b200.shm_handler import (DLROVER_CKPT_CONFIG_KEY, CheckpointConfig,
                                      SharedMemoryHandler, TensorMeta, plan_layout)
from oracle import shm_layout as oracle
from tests.util import meta_to_json

def bit_equal(a, b):
    """tree_equal on the BYTES of the tensors (random bit patterns include NaNs)."""
    if isinstance(a, dict):
        return isinstance(b, dict) and list(a) == list(b) and all(bit_equal(a[k], b[k]) for k in a)
    if isinstance(a, list):
        return isinstance(b, list) and len(a) == len(b) and all(map(bit_equal, a, b))
    if torch.is_tensor(a):
        return (torch.is_tensor(b) and a.dtype == b.dtype and a.shape == b.shape and
                a.contiguous().view(-1).view(torch.uint8).tolist() ==
                b.contiguous().view(-1).view(torch.uint8).tolist()) if a.numel() else \
            (torch.is_tensor(b) and a.dtype == b.dtype and a.shape == b.shape)
    return a == b


DTYPES = [torch.float32, torch.float16, torch.bfloat16, torch.float64, torch.int8, torch.uint8,
          torch.int16, torch.int32, torch.int64, torch.bool]


@st.composite
def tensors(draw):
    dtype = draw(st.sampled_from(DTYPES))
    shape = tuple(draw(st.lists(st.integers(0, 7), min_size=0, max_size=3)))
    seed = draw(st.integers(0, 2**16))
    g = torch.Generator().manual_seed(seed)
    n = int(np.prod(shape)) if shape else 1
    raw = torch.randint(0, 256, (max(n, 1) * 8,), dtype=torch.uint8, generator=g)
    if dtype is torch.bool:
        t = (raw[:n] > 127).reshape(shape)
    else:
        t = raw.view(dtype)[:n].clone().reshape(shape)
    if len(shape) == 2 and shape[0] > 1 and shape[1] > 1 and draw(st.booleans()):
        t = t.t()                       # non-contiguous view
    return t


leaves = st.one_of(tensors(), st.integers(-5, 5), st.text(max_size=4), st.none(),
                   st.tuples(st.integers(0, 3), st.integers(0, 3)), st.booleans())
trees = st.recursive(
    leaves,
    lambda kids: st.one_of(st.lists(kids, max_size=4),
                           st.dictionaries(st.text(min_size=1, max_size=5), kids, max_size=4)),
    max_leaves=14)
state_dicts = st.dictionaries(st.text(min_size=1, max_size=6), trees, min_size=1, max_size=5)


@settings(max_examples=120, deadline=None)
@given(sd=state_dicts)
def test_planner_matches_oracle(sd):
    lay = plan_layout(sd)
    meta, total = oracle.plan_layout(sd)
    assert lay.total == total
    assert meta_to_json(lay.meta, (TensorMeta,)) == meta_to_json(meta, (oracle.OracleTensorMeta,))
    # a warm plan (TensorMeta reuse) is the same plan
    again = plan_layout(sd, lay)
    assert again.total == total and again.meta == lay.meta


@settings(max_examples=40, deadline=None,
          suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(sd=state_dicts)
def test_cpu_save_matches_oracle_image(run_env, sd):
    handler = SharedMemoryHandler(0, host=True)
    try:
        full = dict(sd)
        full[DLROVER_CKPT_CONFIG_KEY] = CheckpointConfig(step=3, paths={})
        handler.save_state_dict(full)
        meta, want = oracle.serialize(sd)       # the config entry holds no tensor
        back = handler.load_state_dict()
        if want.size == 0:
            # nothing to put in a segment: no segment, and (as in the reference,
            # ckpt_saver.py:346-351) nothing to load from memory
            assert handler.shared_memory is None and back == {}
            return
        seg = np.frombuffer(handler.shared_memory.buf, dtype=np.uint8)
        assert seg.size == want.size and np.array_equal(seg, want)
        del seg
        assert back.pop(DLROVER_CKPT_CONFIG_KEY).step == 3
        # what comes back is what the reference's reader makes of the image (0-numel
        # tensors lose their shape: ckpt_saver.py:146-148), not necessarily sd itself
        assert bit_equal(back, oracle.read_image(meta, want))
        del back
    finally:
        handler.unlink()
        handler.close()


@settings(max_examples=60, deadline=None)
@given(sd=state_dicts, threads=st.integers(1, 4))
def test_parallel_writer_equals_torch_save(tmp_path_factory, sd, threads):
    """The agent's parallel torch.save writer produces the file torch.save itself
    writes — byte for byte — for arbitrary nested state dicts (shared storages,
    0-numel tensors, non-contiguous views, non-tensor leaves)."""
    import hashlib

    from dlrover_b200 import fast_torch_save

    d = tmp_path_factory.mktemp("fts")
    # same stem: torch.save stores the file stem as the archive prefix
    a, b = d / "a" / "rank_0.pt", d / "b" / "rank_0.pt"
    a.parent.mkdir()
    b.parent.mkdir()
    # a tensor appearing twice shares its storage in the file
    first = next((v for v in sd.values() if torch.is_tensor(v)), None)
    if first is not None:
        sd = dict(sd, alias_of_first=first)
    torch.save(sd, a)
    fast_torch_save.fast_save(sd, str(b), threads=threads)
    assert hashlib.sha256(a.read_bytes()).hexdigest() == hashlib.sha256(b.read_bytes()).hexdigest()


@settings(max_examples=60, deadline=None)
@given(sizes=st.lists(st.integers(0, 70_000), min_size=1, max_size=12),
       gaps=st.lists(st.integers(0, 17), min_size=12, max_size=12),
       threads=st.integers(1, 6), seed=st.integers(0, 2**16))
def test_host_pack_and_unpack_match_the_oracle(sizes, gaps, threads, seed):
    from dlrover_b200 import _native as native

    rng = np.random.default_rng(seed)
    srcs = [rng.integers(0, 256, size=n, dtype=np.uint8) for n in sizes]
    offs, o = [], gaps[0]
    for i, s in enumerate(srcs):
        offs.append(o)
        o += s.size + gaps[(i + 1) % len(gaps)]
    total = max(o, 1)
    want = oracle.pack_ranges(srcs, offs, total)
    got = np.zeros(total, dtype=np.uint8)
    native.host_pack(got.ctypes.data, [s.ctypes.data if s.size else 0 for s in srcs], offs,
                     [s.size for s in srcs], threads)
    assert np.array_equal(got, want)
    outs = [np.zeros(n, dtype=np.uint8) for n in sizes]
    native.host_unpack(got.ctypes.data, [t.ctypes.data if t.size else 0 for t in outs], offs,
                       list(sizes), threads)
    assert all(np.array_equal(x, y) for x, y in zip(outs, srcs))


# ---------------------------------------------------------------- strided views --

@st.composite
def strided_views(draw):
    """A view with arbitrary slicing (start/stop/step per dim) and an optional permutation of
    a small base tensor — what `w[:, :k]`, `w[::2]`, `w.transpose(0, 1)[3:]` ... produce."""
    ndim = draw(st.integers(1, 4))
    shape = [draw(st.integers(1, 9)) for _ in range(ndim)]
    dtype = draw(st.sampled_from([torch.uint8, torch.int16, torch.float32, torch.float64]))
    n = 1
    for s in shape:
        n *= s
    base = torch.arange(n, dtype=torch.int64).to(dtype).reshape(shape)
    index = []
    for s in shape:
        start = draw(st.integers(0, s - 1))
        stop = draw(st.integers(start + 1, s))
        step = draw(st.integers(1, 3))
        index.append(slice(start, stop, step))
    view = base[tuple(index)]
    if ndim > 1 and draw(st.booleans()):
        view = view.permute(*draw(st.permutations(range(ndim))))
    if draw(st.booleans()):  # a broadcast dim (stride 0): the same rows are read again
        view = view.unsqueeze(0).expand(draw(st.integers(2, 3)), *view.shape)
    return view


@settings(max_examples=300, deadline=None)
@given(view=strided_views(), off=st.integers(0, 1 << 40))
def test_row_ranges_of_any_view_are_the_bytes_copy_would_deposit(view, off):
    """shm_handler._row_ranges either declines (None: the caller repacks the leaf) or
    returns ranges whose concatenation is exactly oracle.tensor_bytes(view) — the bytes the
    reference's `shm_tensor.copy_(view)` leaves (ckpt_saver.py:228-231) — at consecutive
    segment offsets starting at `off`."""
    import ctypes

    from dlrover_b200.shm_handler import _row_ranges

    rows = _row_ranges(view, off, min_row_bytes=1)
    if view.is_contiguous():
        return  # the caller never asks for rows of a dense leaf
    if rows is None:
        return
    want = oracle.tensor_bytes(view).tobytes()
    got = b"".join(ctypes.string_at(p, n) for p, _, n in rows)
    assert got == want
    pos = off
    for _, o, n in rows:
        assert o == pos
        pos += n
    assert pos - off == view.numel() * view.element_size()


# ------------------------------------------------------- cooperative windows --

@settings(max_examples=300, deadline=None)
@given(total=st.integers(0, 1 << 36), n=st.integers(1, 16),
       cuts=st.lists(st.integers(1, 1 << 30), min_size=0, max_size=12))
def test_cooperative_windows_partition_the_image(total, n, cuts):
    """CoopContext.window: the n local ranks' windows are consecutive, 2 MiB aligned (but
    for the end of the image), cover [0, total) exactly once; clipping a list of ranges to
    the windows and putting the pieces together gives the ranges back."""
    from dlrover_b200.shm_handler import CoopContext, _clip_ranges

    wins = [CoopContext(None, i, n, 0).window(total) for i in range(n)]
    assert wins[0][0] == 0 and wins[-1][1] == total
    for (a0, a1), (b0, b1) in zip(wins, wins[1:]):
        assert a1 == b0 and a0 <= a1
        assert a1 % (2 << 20) == 0 or a1 == total
    # ranges laid back to back like a layout, with pseudo pointers
    offs, lens, o = [], [], 0
    for c in cuts:
        if o + c > total:
            break
        offs.append(o)
        lens.append(c)
        o += c
    ptrs = [0x7000_0000_0000 + 4096 * i + (o & 15) for i, o in enumerate(offs)]
    pieces = {}
    for lo, hi in wins:
        for p, o, ln in zip(*_clip_ranges((ptrs, offs, lens), lo, hi)):
            assert lo <= o and o + ln <= hi and ln > 0
            i = max(k for k in range(len(offs)) if offs[k] <= o)
            assert p - ptrs[i] == o - offs[i]            # pointer moved with the offset
            pieces.setdefault(i, []).append((o, ln))
    for i, (o, ln) in enumerate(zip(offs, lens)):
        got = sorted(pieces.get(i, []))
        assert got and got[0][0] == o and sum(x for _, x in got) == ln
        for (a, la), (b, _) in zip(got, got[1:]):
            assert a + la == b
