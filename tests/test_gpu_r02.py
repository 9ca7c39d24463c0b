"""GPU parity tests added in round 2 (all against oracle/shm_layout.py, bit-exact):
  * hybrid saves whose cut is not a multiple of 16 go through the bulk / shift / resid
    tables (arena byte 0 = cut rounded down to 128 B);
  * the staged (bounce-slot) path for host ranges that are not registered — save
    (plain, held, in-place, hybrid) and restore (direct, through the arena);
  * background slice-wise registration followed by plain DMA whose pieces never
    straddle a slice;
  * the ping-pong drain;
  * row-strided leaves through SharedMemoryHandler (save and restore);
  * the hypothesis state-dict generator of test_property_layout.py through the device
    path (every leaf tensor on cuda:0).
"""

import ctypes
import mmap
import os
import time

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings

import _posixshmem
from dlrover_b200 import _native as native
from dlrover_b200.shm_handler import (DLROVER_CKPT_CONFIG_KEY, CheckpointConfig,
                                      SharedMemoryHandler)
from oracle import shm_layout as oracle
from tests.test_property_layout import bit_equal, state_dicts
from tests.util import to_device

pytestmark = pytest.mark.gpu


class _Segment:
    """A plain mmap'd POSIX shm segment (pageable until somebody registers it)."""

    def __init__(self, nbytes):
        self.name = f"/fc_r02_{os.getpid()}_{time.monotonic_ns()}"
        self.fd = _posixshmem.shm_open(self.name, os.O_CREAT | os.O_EXCL | os.O_RDWR, mode=0o600)
        os.ftruncate(self.fd, nbytes)
        self.mm = mmap.mmap(self.fd, nbytes)
        self.addr = ctypes.addressof(ctypes.c_char.from_buffer(self.mm))
        self.nbytes = nbytes

    def bytes(self):
        return np.frombuffer(self.mm, dtype=np.uint8)

    def close(self):
        try:
            self.mm.close()
        except BufferError:
            pass
        os.close(self.fd)
        _posixshmem.shm_unlink(self.name)


def _adamw_like(seed, n_params=5):
    """param, step (4 B!), exp_avg, exp_avg_sq per parameter: everything behind the
    first step scalar is not congruent mod 16 with its (256-B aligned) source."""
    g = torch.Generator().manual_seed(seed)
    leaves = []
    for i in range(n_params):
        n = [300_001, 1 << 20, 4097, (3 << 20) + 5, 70_000][i % 5]
        leaves.append(torch.randint(0, 256, (n,), dtype=torch.uint8, generator=g).cuda())
        leaves.append(torch.randint(0, 256, (4,), dtype=torch.uint8, generator=g).cuda())
        leaves.append(torch.randint(0, 256, (4 * n,), dtype=torch.uint8, generator=g).cuda())
        leaves.append(torch.randint(0, 256, (4 * n,), dtype=torch.uint8, generator=g).cuda())
    offsets, off = [], 0
    for t in leaves:
        offsets.append(off)
        off += t.numel()
    return leaves, offsets, off


@pytest.mark.parametrize("variant", [native.VARIANT_TMA, native.VARIANT_LSU])
def test_hybrid_cut_off_any_alignment_uses_the_same_tables(cuda_device, variant):
    ctx = native.Context(0)
    ctx.set_variant(variant)
    leaves, offsets, total = _adamw_like(3)
    want = oracle.pack_ranges([oracle.tensor_bytes(t) for t in leaves], offsets, total)
    plan = ctx.plan([t.data_ptr() for t in leaves], offsets, [t.numel() for t in leaves])
    host = torch.zeros(total, dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream()
    seen_mod16 = set()
    for first in (1, 2, 3, 6, 9, 14, 19):
        cut = offsets[first]
        seen_mod16.add(cut % 16)
        ctx.arena_reserve(total - (cut & ~127))
        host.zero_()
        k0, _ = ctx.launch_count()
        ticket = plan.save_hybrid_async(host.data_ptr(), cut, stream)
        ctx.save_wait(ticket)
        k1, _ = ctx.launch_count()
        assert 1 <= k1 - k0 <= 3
        assert np.array_equal(host.numpy(), want), (first, cut)
    assert len(seen_mod16) >= 3  # really exercised unaligned cuts
    plan.destroy()
    ctx.destroy()


@pytest.mark.parametrize("threads,slot", [(1, 256 << 10), (3, 1 << 20), (8, 8 << 20)])
def test_staged_path_for_an_unregistered_segment(cuda_device, threads, slot):
    """Nobody pinned the segment: save and restore bounce through the library's pinned
    slots — same image, same restore, for every save flavour."""
    ctx = native.Context(0)
    ctx.set_stage(threads, slot)
    leaves, offsets, total = _adamw_like(7)
    want = oracle.pack_ranges([oracle.tensor_bytes(t) for t in leaves], offsets, total)
    seg = _Segment(total)
    try:
        ctx.arena_reserve(total)
        plan = ctx.plan([t.data_ptr() for t in leaves], offsets, [t.numel() for t in leaves])
        stream = torch.cuda.current_stream()
        cut = offsets[9]
        savers = {
            "snapshot": lambda: plan.save_async(seg.addr, stream),
            "held": lambda: plan.save_async(seg.addr, stream, hold=True),
            "in_place": lambda: plan.save_direct_async(seg.addr, stream),
            "hybrid": lambda: plan.save_hybrid_async(seg.addr, cut, stream),
        }
        for name, save in savers.items():
            seg.bytes()[:] = 0
            ticket = save()
            if name == "held":
                torch.cuda.synchronize()
                assert not seg.bytes().any()
                ctx.save_release(ticket)
            ctx.save_wait(ticket)
            assert ctx.save_pack_done(ticket)
            assert np.array_equal(seg.bytes(), want), name
        keep = [t.clone() for t in leaves]
        for direct in (True, False):
            for t in leaves:
                t.zero_()
            plan.restore_async(seg.addr, stream, direct=direct)
            ctx.restore_wait()
            for t, k in zip(leaves, keep):
                assert torch.equal(t, k), direct
        plan.destroy()
    finally:
        ctx.destroy()
        seg.close()


def test_background_registration_then_plain_dma(cuda_device):
    """fc_host_register_background pins slice by slice; once fc_host_ready says so the
    drain is plain DMA again and no piece straddles a slice (piece size chosen so that
    it would)."""
    ctx = native.Context(0)
    leaves, offsets, total = _adamw_like(11)
    want = oracle.pack_ranges([oracle.tensor_bytes(t) for t in leaves], offsets, total)
    seg = _Segment(total)
    try:
        ctx.arena_reserve(total)
        plan = ctx.plan([t.data_ptr() for t in leaves], offsets, [t.numel() for t in leaves])
        stream = torch.cuda.current_stream()
        # before anything is pinned: staged
        ticket = plan.save_async(seg.addr, stream)
        ctx.save_wait(ticket)
        assert np.array_equal(seg.bytes(), want)
        ctx.host_register_background(seg.addr, total, slice_bytes=2 << 20)
        deadline = time.time() + 60
        while not ctx.host_ready(seg.addr):
            assert time.time() < deadline
            time.sleep(0.01)
        ctx.set_drain(3 << 20, 1)  # 3 MiB pieces over 2 MiB slices
        for mode in (native.DRAIN_HOST_PACED, native.DRAIN_PINGPONG):
            ctx.set_drain_mode(mode)
            seg.bytes()[:] = 0
            _, m0 = ctx.launch_count()
            ticket = plan.save_async(seg.addr, stream)
            ctx.save_wait(ticket)
            _, m1 = ctx.launch_count()
            assert np.array_equal(seg.bytes(), want), mode
            assert m1 - m0 >= total // (2 << 20)  # at least one DMA per slice
        for direct in (True, False):
            keep = [t.clone() for t in leaves]
            for t in leaves:
                t.zero_()
            plan.restore_async(seg.addr, stream, direct=direct)
            ctx.restore_wait()
            for t, k in zip(leaves, keep):
                assert torch.equal(t, k)
        ctx.host_unregister(seg.addr)
        plan.destroy()
    finally:
        ctx.destroy()
        seg.close()


@pytest.mark.parametrize("piece", [64 << 10, 1 << 20, 32 << 20])
def test_pingpong_drain_same_image(cuda_device, piece):
    ctx = native.Context(0)
    ctx.set_drain_mode(native.DRAIN_PINGPONG)
    ctx.set_drain(piece, 1)
    leaves, offsets, total = _adamw_like(13)
    want = oracle.pack_ranges([oracle.tensor_bytes(t) for t in leaves], offsets, total)
    host = torch.zeros(total, dtype=torch.uint8).pin_memory()
    ctx.arena_reserve(total)
    plan = ctx.plan([t.data_ptr() for t in leaves], offsets, [t.numel() for t in leaves])
    stream = torch.cuda.current_stream()
    for hold in (False, True):
        host.zero_()
        ticket = plan.save_async(host.data_ptr(), stream, hold=hold)
        if hold:
            torch.cuda.synchronize()
            assert not host.numpy().any()
            ctx.save_release(ticket)
        ctx.save_wait(ticket)
        assert np.array_equal(host.numpy(), want)
        pack_ms, drain_ms, total_ms = ctx.save_timings(ticket)
        assert drain_ms > 0 and total_ms >= pack_ms
    plan.destroy()
    ctx.destroy()


def test_row_strided_leaves_save_and_restore(cuda_device, run_env):
    """w[:, :k], w[::2], a sliced cube: saved without a device-side .contiguous()
    (one plan range per dense row), restored straight into strided targets."""
    g = torch.Generator().manual_seed(5)
    big = torch.randn(512, 1024, generator=g).cuda()
    cube = torch.randint(-9, 9, (6, 64, 256), dtype=torch.int16, generator=g).cuda()
    sd = {"m": {"cols": big[:, :640], "rows": big[::2], "cube": cube[1:5, ::2, :],
                "dense": big[7], "t": big[:64, :64].t()},   # .t(): no dense rows -> repacked
          "step": 5}
    _, want = oracle.serialize(sd)
    handler = SharedMemoryHandler(0, host=True)
    try:
        full = dict(sd)
        full[DLROVER_CKPT_CONFIG_KEY] = CheckpointConfig(step=2, paths={})
        handler.save_state_dict(full)
        got = np.frombuffer(handler.shared_memory.buf, dtype=np.uint8)
        assert got.size == want.size and np.array_equal(got, want)
        del got
        plan = handler._stager._plans["save"]
        assert plan.n_spans > 512 + 256  # per-row ranges, not 5 repacked tensors
        # restore into fresh strided targets of the same geometry
        big2 = torch.zeros_like(big)
        cube2 = torch.zeros_like(cube)
        tgt = {"m": {"cols": big2[:, :640], "rows": torch.zeros_like(big)[::2],
                     "cube": cube2[1:5, ::2, :], "dense": torch.zeros_like(big[7]),
                     "t": torch.zeros(64, 64, device="cuda")},
               "step": 0}
        handler.restore_into(tgt)
        for k in ("cols", "rows", "cube", "dense", "t"):
            assert torch.equal(tgt["m"][k], sd["m"][k]), k
        assert not big2[:, 640:].any()  # nothing outside the view was touched
    finally:
        handler.unlink()
        handler.close()


@settings(max_examples=25, deadline=None,
          suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(sd=state_dicts)
def test_property_device_path_matches_oracle(cuda_device, run_env, sd):
    """The random nested state dicts of test_property_layout.py with every tensor on
    cuda:0: gather kernels + drain produce the oracle's image, reload and restore_into
    give the bits back."""
    dev_sd = to_device(sd, "cuda")
    handler = SharedMemoryHandler(0, host=True)
    try:
        full = dict(dev_sd)
        full[DLROVER_CKPT_CONFIG_KEY] = CheckpointConfig(step=3, paths={})
        handler.save_state_dict(full)
        meta, want = oracle.serialize(sd)
        if want.size == 0:
            assert handler.shared_memory is None
            return
        seg = np.frombuffer(handler.shared_memory.buf, dtype=np.uint8)
        assert seg.size == want.size and np.array_equal(seg, want)
        del seg
        back = handler.load_state_dict()
        back.pop(DLROVER_CKPT_CONFIG_KEY)
        assert bit_equal(back, oracle.read_image(meta, want))
        del back
    finally:
        handler.unlink()
        handler.close()


def test_user_views_stay_usable_while_the_dma_mapping_is_pinned_in_slices(cuda_device, run_env):
    """The segment is pinned slice by slice through a SECOND mapping; the tensors
    load_state_dict() hands out view the first one, so a plain `param.copy_(view)` of a
    tensor that straddles two slices works (it failed with "invalid argument" when the
    views were on the sliced registration)."""
    sd = {"a": torch.arange(50 << 20, dtype=torch.int32, device="cuda"),      # 200 MiB
          "b": torch.arange(50 << 20, dtype=torch.int32, device="cuda") * 3}  # straddles 256 MiB
    handler = SharedMemoryHandler(0, host=True)
    try:
        full = dict(sd)
        full[DLROVER_CKPT_CONFIG_KEY] = CheckpointConfig(step=1, paths={})
        handler.save_state_dict(full)                 # first save: bounce slots
        assert handler.wait_segment_pinned(60)        # pinned in the background since
        handler.save_state_dict(full)                 # plain DMA now
        views = handler.load_state_dict()
        for k in ("a", "b"):
            target = torch.zeros_like(sd[k])
            target.copy_(views[k])                    # the reference's restore idiom
            assert torch.equal(target, sd[k])
        del views
    finally:
        handler.unlink()
        handler.close()


def test_mapped_plan_compact_arena_scattered_segment_offsets(cuda_device):
    """fc_plan_create_mapped: ranges scattered over a large segment, packed into an arena of
    their own size (arena offset congruent to the source mod 16 -> bulk kernel), drained range
    by range to their segment offsets; restore is the inverse.  What one rank of a "full
    checkpoint from shards" save does."""
    ctx = native.Context(0)
    g = torch.Generator().manual_seed(17)
    sizes = [5, 4096, 300_001, (2 << 20) + 3, 17, 1 << 20]
    leaves = [torch.randint(0, 256, (n,), dtype=torch.uint8, generator=g).cuda() for n in sizes]
    # segment offsets: far apart, odd alignments; holes belong to "other ranks"
    host_offs, o = [], 1000
    for n in sizes:
        host_offs.append(o)
        o += n + 777_777
    total = o
    arena_offs, a = [], 0
    for t in leaves:
        a += (t.data_ptr() - a) % 16
        arena_offs.append(a)
        a += t.numel()
    ctx.arena_reserve(a)
    plan = ctx.plan([t.data_ptr() for t in leaves], arena_offs, sizes, host_offsets=host_offs)
    assert plan.arena_end == a and plan.payload_bytes == sum(sizes)
    host = torch.full((total,), 7, dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream()
    ticket = plan.save_async(host.data_ptr(), stream)
    ctx.save_wait(ticket)
    want = np.full(total, 7, dtype=np.uint8)       # holes untouched
    for t, off in zip(leaves, host_offs):
        want[off:off + t.numel()] = oracle.tensor_bytes(t)
    assert np.array_equal(host.numpy(), want)
    keep = [t.clone() for t in leaves]
    for direct in (False, True):
        for t in leaves:
            t.zero_()
        plan.restore_async(host.data_ptr(), stream, direct=direct)
        ctx.restore_wait()
        for t, k in zip(leaves, keep):
            assert torch.equal(t, k), direct
    # a hybrid save needs the identity mapping
    with pytest.raises(native.NativeError):
        plan.save_hybrid_async(host.data_ptr(), host_offs[2], stream)
    plan.destroy()
    ctx.destroy()
