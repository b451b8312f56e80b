"""Streamed sample store: device ring -> pinned staging -> host arrays, overlapped with the sweeps.

The reference writes every stored iteration into host arrays as it goes (mcmc.py:105-111, sampler.py:89-118).  The
resident device store `[n_iter, C, size]` of the first round caps n_iter by HBM (C3 at 64 chains: 512 MB per stored
iteration) and pays the whole download after the last sweep.  Here the store graph writes slab `it % ring` of a small
device ring (omc_store_copy_ring); after every stored iteration

    compute stream : ... n_thin sweeps | store graph -> slab k         | n_thin sweeps | store graph -> slab k+1 ...
    copy stream    :                      wait(store k) D2H slab k chunks ----------------> event copied[k]
    host threads   :                                     chunk landed -> memcpy into store[it] (several threads)

the compute stream only waits for `copied[k]` before it overwrites slab k again, `ring` stored iterations later.
Host plumbing only: bytes are moved, never computed on.
"""

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

CHUNK_BYTES = 32 << 20
N_STAGING = 8
_pool = {}
_pool_lock = threading.Lock()


def staging_buffers(device_index: int):
    """Process-wide pinned staging buffers of a device (pinning is slow: ~0.5 s per GB on the B200 boxes), created on
    first use -- `warm()` does that on a helper thread while the inputs upload."""
    with _pool_lock:
        if device_index not in _pool:
            _pool[device_index] = [torch.empty(CHUNK_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(N_STAGING)]
        return _pool[device_index]


def host_array(shape) -> np.ndarray:
    """Destination array of a streamed store: np.empty + a transparent-huge-page hint.  The copier threads are the first
    to touch these pages; with 4 KB pages the kernel's fault handling (zeroing, one fault per page) is what bounds them."""
    a = np.empty(shape)
    try:
        import ctypes

        page = 2 << 20
        addr = a.ctypes.data
        lo = (addr + page - 1) // page * page
        hi = (addr + a.nbytes) // page * page
        if hi > lo:
            ctypes.CDLL("libc.so.6", use_errno=True).madvise(ctypes.c_void_p(lo), ctypes.c_size_t(hi - lo), 14)  # MADV_HUGEPAGE
    except Exception:
        pass
    return a


def warm(device_index: int) -> threading.Thread:
    th = threading.Thread(target=staging_buffers, args=(device_index,), daemon=True)
    th.start()
    return th


class StoreStreamer:
    """entries: list of (device ring tensor [ring, ...] float64, host array [n_iter, ...] float64)."""

    def __init__(self, device: torch.device, entries, ring: int, n_iter: int):
        self.device, self.entries, self.ring, self.n_iter = device, entries, int(ring), int(n_iter)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.store_done = [torch.cuda.Event() for _ in range(self.ring)]
        self.copied = [torch.cuda.Event() for _ in range(self.ring)]
        self.issued = [threading.Event() for _ in range(self.n_iter)]
        self.free = queue.Queue()
        for b in staging_buffers(device.index):
            self.free.put(b)
        self.q = queue.Queue()
        self.pool = ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 4) // 2)))
        self.futures = []
        self.error = None
        self.d2h_bytes = 0
        self.thread = threading.Thread(target=self._drain, daemon=True)
        self.thread.start()

    # ---- called by the thread that launches the sweeps
    def before_store(self, it: int, stream: torch.cuda.Stream):
        """Slab it % ring is about to be overwritten: its previous contents (iteration it - ring) must have left."""
        if it >= self.ring:
            self.issued[it - self.ring].wait()
            self._check()
            stream.wait_event(self.copied[it % self.ring])

    def after_store(self, it: int, stream: torch.cuda.Stream):
        self.store_done[it % self.ring].record(stream)
        self.q.put(it)

    def finish(self):
        self.q.put(None)
        self.thread.join()
        for f in self.futures:
            f.result()
        self.pool.shutdown()
        self._check()

    def _check(self):
        if self.error is not None:
            raise self.error

    # ---- drain thread: issues the device->host copies of one slab after the other, hands landed chunks to the copiers
    def _drain(self):
        try:
            torch.cuda.set_device(self.device)
            while True:
                it = self.q.get()
                if it is None:
                    return
                slot = it % self.ring
                self.copy_stream.wait_event(self.store_done[slot])
                for ring_t, host in self.entries:
                    src = ring_t[slot].reshape(-1).view(torch.uint8)
                    dst = host[it].reshape(-1).view(np.uint8)
                    nbytes = src.numel()
                    self.d2h_bytes += nbytes
                    for off in range(0, nbytes, CHUNK_BYTES):
                        cnt = min(CHUNK_BYTES, nbytes - off)
                        buf = self.free.get()
                        with torch.cuda.stream(self.copy_stream):
                            buf[:cnt].copy_(src[off:off + cnt], non_blocking=True)
                            ev = torch.cuda.Event()
                            ev.record(self.copy_stream)
                        self.futures.append(self.pool.submit(self._land, ev, buf, cnt, dst[off:off + cnt]))
                self.copied[slot].record(self.copy_stream)
                self.issued[it].set()
        except BaseException as exc:   # surface in the launching thread instead of dying silently
            self.error = exc
            for e in self.issued:
                e.set()

    def _land(self, ev, buf, cnt, dst):
        try:
            ev.synchronize()
            np.copyto(dst, buf[:cnt].numpy())
        finally:
            self.free.put(buf)
