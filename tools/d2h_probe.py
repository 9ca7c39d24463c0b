"""Does a small D2H copy wait for a large in-flight D2H on another stream?
pageable (.item()) vs pinned destination; big copy as one piece vs paced pieces."""
import json
import time

import torch

torch.cuda.set_device(0)
big = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
host = torch.empty(4 << 30, dtype=torch.uint8).pin_memory()
small = torch.ones(4, device="cuda")
pin_small = torch.empty(4).pin_memory()
side = torch.cuda.Stream()
main = torch.cuda.current_stream()
res = {}


def big_copy(piece):
    with torch.cuda.stream(side):
        if piece is None:
            host.copy_(big, non_blocking=True)
        else:
            for o in range(0, big.numel(), piece):
                host[o:o + piece].copy_(big[o:o + piece], non_blocking=True)


for name, piece in (("one_piece", None), ("pieces_32M", 32 << 20), ("pieces_4M", 4 << 20)):
    for kind in ("item_pageable", "pinned_async", "none"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        big_copy(piece)
        t1 = time.perf_counter()
        if kind == "item_pageable":
            small[0].item()
        elif kind == "pinned_async":
            pin_small.copy_(small, non_blocking=True)
            main.synchronize()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        res[f"{name}/{kind}"] = {"submit_ms": round((t1 - t0) * 1e3, 2),
                                 "small_copy_ms": round((t2 - t1) * 1e3, 2),
                                 "total_ms": round((t3 - t0) * 1e3, 2)}
print(json.dumps(res, indent=1))
