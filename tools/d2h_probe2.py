"""Paced big D2H (<=2 pieces in flight, submitted from a helper thread) while the
main thread issues small D2H copies: how long does each small copy take?"""
import json
import threading
import time

import torch

torch.cuda.set_device(0)
big = torch.empty(8 << 30, dtype=torch.uint8, device="cuda")
host = torch.empty(8 << 30, dtype=torch.uint8).pin_memory()
small = torch.ones(4, device="cuda")
pin_small = torch.empty(4).pin_memory()
side = torch.cuda.Stream()
main = torch.cuda.current_stream()


def paced(piece, depth):
    evs = []
    with torch.cuda.stream(side):
        k = 0
        for o in range(0, big.numel(), piece):
            if k >= depth:
                evs[k - depth].synchronize()
            host[o:o + piece].copy_(big[o:o + piece], non_blocking=True)
            e = torch.cuda.Event()
            e.record(side)
            evs.append(e)
            k += 1
    side.synchronize()


out = {}
for piece, depth in ((32 << 20, 2), (8 << 20, 2), (32 << 20, 1), (256 << 20, 2)):
    for kind in ("item", "pinned"):
        torch.cuda.synchronize()
        th = threading.Thread(target=paced, args=(piece, depth))
        t0 = time.perf_counter()
        th.start()
        lat = []
        while th.is_alive():
            a = time.perf_counter()
            if kind == "item":
                small[0].item()
            else:
                pin_small.copy_(small, non_blocking=True)
                main.synchronize()
            lat.append((time.perf_counter() - a) * 1e3)
            time.sleep(0.005)
        th.join()
        total = (time.perf_counter() - t0) * 1e3
        lat.sort()
        out[f"piece{piece >> 20}M_depth{depth}/{kind}"] = {
            "big_total_ms": round(total, 1), "big_GBps": round(big.numel() / total / 1e6, 1),
            "n_small": len(lat), "small_med_ms": round(lat[len(lat) // 2], 3),
            "small_max_ms": round(lat[-1], 3)}
print(json.dumps(out, indent=1))
