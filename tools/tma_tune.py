"""Aligned gather kernel: launch-shape sweep (run with FC_TMA_HINTS=0..3)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlrover_b200 import _native as native
from dlrover_b200 import shapes

torch.cuda.set_device(0)
ctx = native.get_context(0)
sd = shapes.build_state_dict(shapes.llama3_8b_shapes(), torch.bfloat16, "cuda", fill=False)
leaves = list(sd.values())
offs, o = [], 0
for t in leaves:
    offs.append(o)
    o += t.numel() * 2
ctx.arena_reserve(o)
s = torch.cuda.current_stream()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
CONFIGS = []
for tile_k in (64, 80, 96, 104, 112):
    for chunk_mult in (2, 3, 4, 8):
        CONFIGS.append((tile_k * chunk_mult << 10, 1, 2, tile_k << 10))
for tile_k in (40, 48, 56):
    for chunk_mult in (4, 8):
        CONFIGS.append((tile_k * chunk_mult << 10, 2, 2, tile_k << 10))
CONFIGS += [(256 << 10, 1, 2, 96 << 10), (256 << 10, 2, 3, 32 << 10), (288 << 10, 1, 3, 72 << 10)]
plans = {}
for chunk, cps, st, tile in CONFIGS:
    if (st * tile + 8 * st) > (227 << 10):
        continue
    if chunk not in plans:
        plans[chunk] = ctx.plan([t.data_ptr() for t in leaves], offs,
                                [t.numel() * 2 for t in leaves], chunk)
    plan = plans[chunk]
    ctx.set_launch(tma_ctas_per_sm=cps, tma_stages=st, tma_tile_bytes=tile)
    for _ in range(3):
        plan.pack(s, native.VARIANT_TMA)
    a.record()
    for _ in range(6):
        plan.pack(s, native.VARIANT_TMA)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 6
    print(f"cfg chunk={chunk>>10}K ctas={cps} stages={st} tile={tile>>10}K ms={ms:.3f} "
          f"GB/s={2*o/ms/1e6:.0f}", flush=True)
