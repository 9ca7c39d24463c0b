"""Small driver for compute-sanitizer: every kernel variant on ragged,
misaligned ranges (pack + unpack), checked against numpy."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlrover_b200 import _native as native

torch.cuda.set_device(0)
ctx = native.get_context(0)
g = torch.Generator().manual_seed(3)
base = torch.randint(0, 256, (24 << 20,), dtype=torch.uint8, generator=g).cuda()
tensors, offs, o = [], [], 5
cursor = 0
for i, n in enumerate([1, 17, 4095, 4096, 70_001, 1 << 20, (3 << 20) + 7, 300_000, 2 << 20, 33]):
    start = cursor + (i * 7) % 16 + 1  # every source alignment class
    tensors.append(base[start:start + n])
    cursor = start + n + 64
    offs.append(o)
    o += n + (i % 3)
ctx.arena_reserve(o)
plan = ctx.plan([t.data_ptr() for t in tensors], offs, [t.numel() for t in tensors], 64 << 10)
host = torch.zeros(o, dtype=torch.uint8).pin_memory()
for variant in (native.VARIANT_TMA, native.VARIANT_LSU):
    ctx.set_variant(variant)
    tk = plan.save_async(host.data_ptr(), torch.cuda.current_stream())
    ctx.save_wait(tk)
    img = host.numpy()
    for t, off in zip(tensors, offs):
        assert np.array_equal(img[off:off + t.numel()], t.cpu().numpy())
    keep = [t.clone() for t in tensors]
    for t in tensors:
        t.zero_()
    plan.restore_async(host.data_ptr(), torch.cuda.current_stream())
    ctx.restore_wait()
    for a, b in zip(tensors, keep):
        assert torch.equal(a, b)
# hybrid / in-place saves: the bulk / shift / resid table slices from the cut on, into an
# arena whose byte 0 stands for the cut rounded down to 128 B; in-place restore
ctx.set_variant(native.VARIANT_AUTO)
for first in (0, 3, 6, len(tensors)):
    cut = offs[first] if first < len(tensors) else o
    host.zero_()
    tk = plan.save_hybrid_async(host.data_ptr(), cut, torch.cuda.current_stream())
    ctx.save_sources_wait(tk)
    ctx.save_wait(tk)
    img = host.numpy()
    for t, off in zip(tensors, offs):
        assert np.array_equal(img[off:off + t.numel()], t.cpu().numpy())
keep = [t.clone() for t in tensors]
for t in tensors:
    t.zero_()
plan.restore_async(host.data_ptr(), torch.cuda.current_stream(), direct=True)
ctx.restore_wait()
for a, b in zip(tensors, keep):
    assert torch.equal(a, b)
# staged path (host range nobody pinned) and the ping-pong drain
pageable = np.zeros(o, dtype=np.uint8)
ctx.set_stage(3, 256 << 10)
tk = plan.save_async(pageable.ctypes.data, torch.cuda.current_stream())
ctx.save_wait(tk)
for t, off in zip(tensors, offs):
    assert np.array_equal(pageable[off:off + t.numel()], t.cpu().numpy())
for t in tensors:
    t.zero_()
plan.restore_async(pageable.ctypes.data, torch.cuda.current_stream())
ctx.restore_wait()
for a, b in zip(tensors, keep):
    assert torch.equal(a, b)
ctx.set_drain_mode(native.DRAIN_PINGPONG)
ctx.set_drain(128 << 10, 1)
host.zero_()
tk = plan.save_async(host.data_ptr(), torch.cuda.current_stream())
ctx.save_wait(tk)
for t, off in zip(tensors, offs):
    assert np.array_equal(host.numpy()[off:off + t.numel()], t.cpu().numpy())
print("SANITIZE_TARGET_OK")
