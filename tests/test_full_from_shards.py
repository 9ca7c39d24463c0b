"""Host-side pieces of "the FULL checkpoint assembled from shards" (no gather):
_box_ranges / plan_layout_full_from_shards / _clip_triples (shm_handler.py) and the
NoShardData container of the FSDP engine."""

import ctypes
import pickle

import numpy as np
import pytest
import torch

from dlrover_b200.flash_checkpoint.fsdp_engine import NoShardData
from dlrover_b200.shm_handler import (_box_ranges, _clip_triples, plan_layout,
                                      plan_layout_full_from_shards)
from oracle import shm_layout as oracle


def _paint(image, ranges):
    for piece, off, n in ranges:
        assert piece.is_contiguous() and piece.numel() * piece.element_size() == n
        image[off:off + n] = np.frombuffer(ctypes.string_at(piece.data_ptr(), n), dtype=np.uint8)


@pytest.mark.parametrize("shape,box_off,box_shape", [
    ((12, 8), (3, 0), (4, 8)),            # dim-0 shard: one contiguous range
    ((12, 8), (0, 2), (12, 3)),           # column shard: one range per row
    ((4, 6, 10), (1, 0, 0), (2, 6, 10)),  # dim-0 shard of a cube
    ((4, 6, 10), (0, 2, 0), (4, 3, 10)),  # dim-1 shard: one range per leading index
    ((4, 6, 10), (1, 1, 4), (2, 3, 5)),   # general box
    ((7,), (2,), (3,)),
    ((), (), ()),
])
def test_box_ranges_paint_the_box_into_the_full_tensor(shape, box_off, box_shape):
    full = torch.arange(int(np.prod(shape)) if shape else 1, dtype=torch.int32).reshape(shape)
    index = tuple(slice(o, o + s) for o, s in zip(box_off, box_shape))
    local = full[index].clone() if shape else full.clone()
    image = np.zeros(100 + full.numel() * 4, dtype=np.uint8)
    _paint(image, _box_ranges(local, box_off, shape, full_off=100))
    want = np.zeros_like(image)
    masked = torch.zeros_like(full)
    if shape:
        masked[index] = full[index]
    else:
        masked = full.clone()
    want[100:] = oracle.tensor_bytes(masked)
    assert np.array_equal(image, want)


class _FakeSharded:
    """Duck-types a ShardedTensor: size(), dtype, local_shards()."""

    class _Shard:
        def __init__(self, tensor, offsets):
            self.tensor = tensor
            self.metadata = type("M", (), {"shard_offsets": list(offsets),
                                           "shard_sizes": list(tensor.shape)})()

    def __init__(self, full, rows):
        self._full, self.dtype = full, full.dtype
        lo, hi = rows
        self._shards = [self._Shard(full[lo:hi].clone(), (lo,) + (0,) * (full.dim() - 1))] \
            if hi > lo else []

    def size(self):
        return self._full.size()

    def local_shards(self):
        return self._shards


def test_two_ranks_shards_tile_the_full_image():
    """Every rank plans the SAME full layout from its own shards; the union of the ranks'
    ranges + the replicated leaves is the oracle image of the gathered state dict."""
    g = torch.Generator().manual_seed(9)
    full = {"model": {"w": torch.randn(10, 6, generator=g), "b": torch.randn(7, generator=g),
                      "e": torch.randn(3, 5, generator=g).to(torch.bfloat16)},
            "optim": {"state": {"w": {"step": torch.tensor(4.0),
                                      "exp_avg": torch.randn(10, 6, generator=g)}},
                      "param_groups": [{"lr": 0.1, "params": ["w", "b", "e"]}]},
            "epoch": 3}
    rows = {"w": [(0, 5), (5, 10)], "b": [(0, 4), (4, 7)], "e": [(0, 2), (2, 3)]}

    def view(rank):
        def shard(name, t):
            return _FakeSharded(t, rows[name][rank])
        return {"model": {k: shard(k, v) for k, v in full["model"].items()},
                "optim": {"state": {"w": {"step": full["optim"]["state"]["w"]["step"],
                                          "exp_avg": shard("w", full["optim"]["state"]["w"]["exp_avg"])}},
                          "param_groups": full["optim"]["param_groups"]},
                "epoch": 3}

    _, want = oracle.serialize(full)
    ref_layout = plan_layout(full)
    image = np.zeros(want.size, dtype=np.uint8)
    for rank in (0, 1):
        lay, mine = plan_layout_full_from_shards(view(rank))
        assert lay.total == ref_layout.total == want.size
        assert [(m.shape, m.dtype, m.offset) for m in lay.leaf_metas] == \
               [(m.shape, m.dtype, m.offset) for m in ref_layout.leaf_metas]
        _paint(image, mine)
        if rank == 0:   # the replicated leaf (step): written once
            _paint(image, [(t, m.offset, m.numel * m.element_size) for t, m in lay.host_leaves])
    assert np.array_equal(image, want)


def test_clip_triples_cuts_tensors_at_window_borders():
    a = torch.arange(100, dtype=torch.int16)       # bytes [10, 210)
    b = torch.arange(50, dtype=torch.int64)        # bytes [210, 610)
    triples = [(a, 10, 200), (b, 210, 400)]
    assert _clip_triples(triples, 0, 1000) == triples
    cut = _clip_triples(triples, 100, 300)
    assert [(o, n) for _, o, n in cut] == [(100, 110), (210, 90)]
    assert bytes(cut[0][0].numpy()) == a.view(torch.uint8)[90:200].numpy().tobytes()
    assert bytes(cut[1][0].numpy()) == b.view(torch.uint8)[:90].numpy().tobytes()
    assert _clip_triples(triples, 700, 800) == []


def test_no_shard_data_round_trip():
    steps = {f"state.p{i}.step": torch.tensor(float(i)) for i in range(5)}
    blob, index, off = [], {}, 0
    for k, t in steps.items():
        raw = t.numpy().tobytes()
        index[k] = (t.dtype, tuple(t.shape), off, len(raw))
        blob.append(raw)
        off += len(raw)
    data = NoShardData({"param_groups": [{"lr": 0.1}], "name": "x"}, b"".join(blob), index)
    back = pickle.loads(pickle.dumps(data))
    assert sorted(back.keys()) == sorted(list(steps) + ["param_groups", "name"])
    assert len(back) == 7 and "state.p3.step" in back and "nope" not in back
    assert float(back["state.p3.step"]) == 3.0 and back["state.p3.step"].shape == ()
    assert back["param_groups"] == [{"lr": 0.1}] and back.get("missing", 7) == 7
    assert len(pickle.dumps(data)) < 1200   # a handful of tensors pickled by torch: ~10x that


from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@st.composite
def _grids(draw):
    """A tensor shape and a grid partition of it (cut points per dim) — DTensor Shard(d) on a
    mesh, 2-D HSDP x TP style boxes, uneven last shards included."""
    ndim = draw(st.integers(1, 3))
    shape = [draw(st.integers(1, 12)) for _ in range(ndim)]
    cuts = []
    for s in shape:
        k = draw(st.integers(0, min(3, s - 1)))
        inner = sorted(draw(st.lists(st.integers(1, s - 1), min_size=k, max_size=k, unique=True))) \
            if s > 1 else []
        cuts.append([0] + inner + [s])
    dtype = draw(st.sampled_from([torch.uint8, torch.bfloat16, torch.float32, torch.int64]))
    return tuple(shape), cuts, dtype


@settings(max_examples=200, deadline=None)
@given(grid=_grids(), full_off=st.integers(0, 4096))
def test_any_grid_of_boxes_tiles_the_full_tensor(grid, full_off):
    """_box_ranges: painting every box of a grid partition at its place gives the bytes of the
    gathered tensor (oracle.tensor_bytes) — each byte written exactly once."""
    import itertools

    shape, cuts, dtype = grid
    n = int(np.prod(shape))
    full = torch.arange(n, dtype=torch.int64).to(dtype).reshape(shape) if dtype != torch.bfloat16 \
        else (torch.arange(n, dtype=torch.float32) % 251).to(torch.bfloat16).reshape(shape)
    es = full.element_size()
    image = np.zeros(full_off + n * es, dtype=np.uint8)
    hits = np.zeros(full_off + n * es, dtype=np.uint8)
    for corner in itertools.product(*[range(len(c) - 1) for c in cuts]):
        lo = [cuts[d][i] for d, i in enumerate(corner)]
        hi = [cuts[d][i + 1] for d, i in enumerate(corner)]
        local = full[tuple(slice(a, b) for a, b in zip(lo, hi))].clone()
        ranges = _box_ranges(local, tuple(lo), shape, full_off=full_off)
        _paint(image, ranges)
        for _, off, nb in ranges:
            hits[off:off + nb] += 1
    assert np.array_equal(image[full_off:], oracle.tensor_bytes(full))
    assert bool((hits[full_off:] == 1).all()) and not image[:full_off].any()
