"""Tuning aid: ways of bringing a large device store back to the host (what MCMC.collect does), timed."""
import time

import numpy as np
import torch

gb = 3.0
n = int(gb * 2**30 / 8)
d = torch.randn(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()


def t(label, fn):
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{label:50s} {dt:7.3f} s  {gb / dt:6.2f} GB/s", flush=True)
    return out


t("pageable  d.cpu()", lambda: d.cpu())
t("pageable  d.cpu() again", lambda: d.cpu())
h = t("pinned alloc torch.empty(pin_memory=True)", lambda: torch.empty(n, dtype=torch.float64, pin_memory=True))
t("D2H into pinned (non_blocking)", lambda: h.copy_(d, non_blocking=True))
t("D2H into pinned again", lambda: h.copy_(d, non_blocking=True))
del h
stage = [torch.empty(32 * 2**20 // 8, dtype=torch.float64, pin_memory=True) for _ in range(2)]


def staged():
    out = np.empty(n, dtype=np.float64)
    ot = torch.from_numpy(out)
    m = stage[0].numel()
    ev = [torch.cuda.Event(), torch.cuda.Event()]
    k = 0
    pending = []
    for off in range(0, n, m):
        b = k & 1
        if len(pending) >= 2:
            po, pb, pe, cnt = pending.pop(0)
            pe.synchronize()
            ot[po:po + cnt].copy_(stage[pb][:cnt])
        cnt = min(m, n - off)
        stage[b][:cnt].copy_(d[off:off + cnt], non_blocking=True)
        ev[b].record()
        pending.append((off, b, ev[b], cnt))
        k += 1
    for po, pb, pe, cnt in pending:
        pe.synchronize()
        ot[po:po + cnt].copy_(stage[pb][:cnt])
    return out


t("staged through 2 x 32 MB pinned into np.empty", staged)
t("staged again", staged)
t("np.empty + first touch only (fill 0)", lambda: np.zeros(n) + 0)
import ctypes
t("cudaHostRegister of np.empty(3GB)", lambda: torch.cuda.cudart().cudaHostRegister(np.empty(n).ctypes.data, n * 8, 0))
