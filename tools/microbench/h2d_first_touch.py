"""Is the first host->device DMA out of a fresh pinned buffer slower than the following ones?  (The C2 e2e leg uploads
21 GB once; on a fresh box that upload ran at 14-25 GB/s, in later processes at 58 GB/s.)"""
import time

import torch

torch.cuda.init()
n = 8 << 30
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
for how in ("cpu_fill", "d2h_fill"):
    t0 = time.perf_counter()
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    t1 = time.perf_counter()
    if how == "cpu_fill":
        h.fill_(1)
    else:
        h.copy_(dst)
        torch.cuda.synchronize()
    t2 = time.perf_counter()
    rates = []
    for _ in range(4):
        t = time.perf_counter()
        dst.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        rates.append(n / (time.perf_counter() - t) / 1e9)
    print(how, "alloc %.2f s fill %.2f s" % (t1 - t0, t2 - t1), "H2D GB/s per pass:", ["%.1f" % r for r in rates])
    del h
