"""ncu target: gather kernel on the Llama-3-8B state_dict with every arena
offset displaced by 4 bytes (source/destination not congruent mod 16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlrover_b200 import _native as native
from dlrover_b200 import shapes

torch.cuda.set_device(0)
ctx = native.get_context(0)
shift = int(os.getenv("SHIFT", "4"))
sd = shapes.build_state_dict(shapes.llama3_8b_shapes(), torch.bfloat16, "cuda", fill=False)
leaves = list(sd.values())
offs, o = [], shift
for t in leaves:
    offs.append(o)
    o += t.numel() * 2
ctx.arena_reserve(o + 64)
plan = ctx.plan([t.data_ptr() for t in leaves], offs, [t.numel() * 2 for t in leaves])
s = torch.cuda.current_stream()
for _ in range(4):
    plan.pack(s, native.VARIANT_AUTO)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    plan.pack(s, native.VARIANT_AUTO)
b.record()
torch.cuda.synchronize()
print("shift", shift, "ms/launch", a.elapsed_time(b) / 5, "GB/s", 2 * (o - shift) / (a.elapsed_time(b) / 5) / 1e6)
if os.getenv("SWEEP"):
    for cps in (1, 2, 3, 4, 6, 8):
        ctx.set_launch(lsu_ctas_per_sm=cps)
        for _ in range(2):
            plan.pack(s, native.VARIANT_AUTO)
        a.record()
        for _ in range(5):
            plan.pack(s, native.VARIANT_AUTO)
        b.record()
        torch.cuda.synchronize()
        print("ctas/sm", cps, "ms", round(a.elapsed_time(b) / 5, 3))

if os.getenv("SWEEP_SHIFT"):
    for cps, st, tile in ((2, 3, 16), (2, 2, 24), (1, 2, 32), (1, 2, 48), (1, 3, 32), (1, 2, 56),
                          (2, 2, 16), (1, 4, 24), (3, 2, 16), (1, 2, 40)):
        try:
            ctx.set_shift_launch(cps, st, tile << 10)
        except native.NativeError as e:
            print("skip", cps, st, tile, e)
            continue
        for _ in range(2):
            plan.pack(s, native.VARIANT_AUTO)
        a.record()
        for _ in range(5):
            plan.pack(s, native.VARIANT_AUTO)
        b.record()
        torch.cuda.synchronize()
        print("shiftcfg ctas", cps, "stages", st, "tileK", tile, "ms", round(a.elapsed_time(b) / 5, 3))
